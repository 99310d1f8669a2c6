"""BASELINE configs[3]: Plummer N=1M with 10 % primordial binaries (test_multiples2.py:236-304) + neighbour lists,
j sharded over the ranks (one per GPU).  Run under torchrun (any world size) or plain python (1 GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4.py [N]

Measures (i) the full sweep of all 1.1M i through the fused peer-memory exchange, (ii) parity of a sampled
i-subset (binary members included) against the FP64 oracle on rank 0 and Newton's third law over the whole sweep,
(iii) neighbour lists (h2 = min(8 dnn^2, 1), gpu.cc:629-630) of a sampled i-block through the g6 ABI on every
shard, merged over the ranks and compared with the oracle, (iv) on 1 GPU only: the unmodified ph4 integrator
(-DGPU objects on this library) for a bounded number of block steps -> seconds per block step and, by
extrapolation, per N-body time unit (ph4 itself is single-process without MPI)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from amuse_b200 import g6lib, plummer as P, sharding as S  # noqa: E402


def main():
    n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    ph4_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m, x, v = P.new_plummer_model(n0, seed=1)
    ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1)
    n = len(m)
    j0, j1 = S.define_domain(n, world, rank)
    g = g6lib.G6(local)
    L = g.L
    L.g6x_set_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), 1)
    L.g6x_set_j_offset(j0)
    g.set_j_particles(ids[j0:j1], m[j0:j1], x[j0:j1], v[j0:j1])
    njl = j1 - j0
    if world > 1:
        S.attach_peers(L, n)
    d_id = torch.from_numpy(ids).to(dev); d_x = torch.from_numpy(x).to(dev); d_v = torch.from_numpy(v).to(dev)
    d_sum = torch.zeros((n, 7), dtype=torch.float64, device=dev)
    d_key = torch.zeros(n, dtype=torch.int64, device=dev)
    d_nn = torch.zeros(n, dtype=torch.int32, device=dev)
    calc = L.g6x_calc_device_allreduce if world > 1 else L.g6x_calc_device

    def sweep(t):
        L.g6x_predict(njl, float(t))
        calc(njl, n, d_id.data_ptr(), d_x.data_ptr(), d_v.data_ptr(), None, 0.0, 1, d_sum.data_ptr(), d_key.data_ptr(),
             d_nn.data_ptr())

    sweep(0.0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 2
    for k in range(reps):
        sweep(1e-9 * (k + 1))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sweep(0.0)
    torch.cuda.synchronize()
    res = {"config": "Plummer N=%d + 10%% binaries = %d particles, eps2=0, j sharded over %d GPU(s)" % (n0, n, world),
           "max_id": int(ids.max()), "sweep_ms": float(ms.item()),
           "interactions_per_s": float(n) * n / (ms.item() * 1e-3)}
    s = d_sum.cpu().numpy(); nn = d_nn.cpu().numpy()
    acc, jerk = s[:, 0:3], s[:, 3:6]
    res["newton3_acc"] = float(np.abs((m[:, None] * acc).sum(axis=0)).max() / (m * np.linalg.norm(acc, axis=1)).sum())
    res["newton3_jerk"] = float(np.abs((m[:, None] * jerk).sum(axis=0)).max() / (m * np.linalg.norm(jerk, axis=1)).sum())
    # binary members must find their companion as nearest neighbour
    nbin = n - n0
    prim = np.arange(0, n0, n0 // nbin)[:nbin]
    res["binaries_mutual_nn"] = float(np.mean((nn[prim] == ids[n0:]) & (nn[n0:] == ids[prim])))

    # ---- sampled parity against the FP64 oracle (rank 0) ----------------------------------------
    rnd = np.random.RandomState(0)
    samp = np.sort(np.concatenate([rnd.choice(n0, 96, replace=False), prim[:16], n0 + np.arange(16)]))
    if rank == 0:
        from oracle import oracle as O
        from helpers import check_forces, check_nn, error_report
        ref = O.force(x[samp], v[samp], m, x, v, 0.0, iid=ids[samp], jid=ids, scales=True)
        got = dict(acc=acc[samp], jerk=jerk[samp], pot=-s[samp, 6])
        res["parity_sample"] = error_report(got, ref)
        try:
            check_forces(got, ref, what="config 4 sample")
            check_nn(nn[samp], ref["nn"], ids, x[samp], x)
            res["parity_ok"] = True
        except AssertionError as e:
            res["parity_ok"] = False
            res["parity_msg"] = str(e)
        dnn_s = ref["dnn"]
    else:
        dnn_s = np.zeros(len(samp))
    if world > 1:
        t_ = torch.from_numpy(dnn_s).to(dev)
        dist.broadcast(t_, 0)
        dnn_s = t_.cpu().numpy()

    # ---- neighbour lists of the sampled block through the g6 ABI on every shard, merged over ranks ----
    L.g6x_set_stream(None, 0)
    h2 = np.minimum(8 * dnn_s ** 2, 1.0)
    g.set_ti(0.0)
    t0 = time.perf_counter()
    g.calc(ids[samp], x[samp], v[samp], 0.0, h2=h2)
    overflow = g.read_neighbour_list()
    cnt, lst = [], []
    for i in range(len(samp)):
        rc, c, l = g.get_neighbour_list(i)
        cnt.append(c); lst.append(l.tolist())
    tot, merged = S.gather_neighbour_lists(cnt, lst)
    res["ngb_seconds"] = time.perf_counter() - t0
    res["ngb_overflow"] = int(overflow)
    res["ngb_mean_len"] = float(np.mean(tot))
    if rank == 0:
        bad = 0
        for k, i in enumerate(samp):
            c, l = O.neighbours(int(ids[i]), x[i], h2[k], ids, m, x)
            r2 = ((x - x[i]) ** 2).sum(axis=1)
            edge = set(ids[np.abs(r2 - h2[k]) <= 1e-6 * h2[k]].tolist())
            if not (set(merged[k]) ^ set(l.tolist()) <= edge):
                bad += 1
        res["ngb_lists_wrong"] = bad
    g.close()

    # ---- the unmodified ph4 integrator on this library, bounded number of block steps (1 GPU) ----
    if world == 1 and ph4_steps > 0:
        from oracle import oracle as O
        if O.ref_available("libph4ref_gpu.so"):
            r = O.ref_evolve(m, x, v, 0.0, 0.14, 1.0, ids=ids, use_gpu=True, libname="libph4ref_gpu.so",
                             max_block_steps=ph4_steps)
            res["ph4"] = {"block_steps": r["block_steps"], "particle_steps": r["particle_steps"], "t": r["t"],
                          "seconds": r["seconds"], "ms_per_block_step": 1e3 * r["seconds"] / max(1, r["block_steps"]),
                          "s_per_nbody_unit_extrapolated": r["seconds"] / r["t"] if r["t"] > 0 else None,
                          "dE_over_E": abs((r["E1"] - r["E0"]) / r["E0"]),
                          "note": "no encounter management (standalone ph4, eps2=0): hard binaries set the step"}
    if rank == 0:
        print("CONFIG4 " + json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
