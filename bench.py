#!/usr/bin/env python
"""Benchmark of the g6 Hermite force path (BASELINE.json metric: interactions/s and % of FP32
peak for full i-block force sweeps on synthetic Plummer spheres, headline N = 1M).

    python bench.py --gpus N --steps K --warmup W            # this library on N B200s
    python bench.py --impl reference --gpus N ...            # the reference CPU path (rank 0)

One "step" = one full i-block force sweep: predict all j to t, then acc/jerk/pot/nearest
neighbour of all N i-particles against all N j-particles (N^2 interactions).

* value : device-timed (CUDA events on the launching stream), inputs resident in HBM,
          j sharded over the ranks, partial forces combined by an NCCL all-reduce
          (sum acc/jerk/pot, min nearest-neighbour key, sum of resolved ids).
* e2e   : N=1: the same sweep through the g6 C ABI with HOST buffers (g6_set_ti_, then
          g6calc_firsthalf_/g6calc_lasthalf2_ per 16384-particle chunk; H2D/D2H inside).
          N>1: pinned host inputs -> H2D -> device entry point -> all-reduce -> D2H.
* roofline : FP32 pipe, 60 flop/interaction convention (src/amuse_ph4/src/jdata.cc:1038).
* cpu_baseline : the reference's own double-precision force loop (oracle/_ref, built from the
          unmodified ph4 sources) or the oracle port, timed on this box's host cores on a bounded
          i-sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 60.0          # convention, src/amuse_ph4/src/jdata.cc:1038
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # SMs x FP32 lanes x 2 x clocks.max.sm (B200_PROFILING.md)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1 << 20,
                    help="particles (default 1M, the headline size); spell it --particles under torchrun, whose "
                         "own parser treats --n as an abbreviation")
    ap.add_argument("--eps2", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--shuffle-ids", action="store_true", help="ids permuted against the addresses")
    ap.add_argument("--parity-sample", type=int, default=1024,
                    help="i-particles checked against the oracle after the timed region (0: skip)")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the block-step latency table and the bounded ph4 run (N=1 only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: combine the j-shard partials inside the library over peer memory (default) or with "
                         "three NCCL all-reduces (the reference implementation of the exchange)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm (reference): ph4's own force loop on the host cores.
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Seconds spent in the reference force loop itself (setup of the reference's jdata excluded)."""
    kind, mass, pos, vel, eps2, lo, hi = args
    from oracle import oracle as O
    if kind == "reference":
        n = len(mass)
        z = np.zeros((n, 3))
        r = O.ref_predict_force(mass, np.zeros(n), pos, vel, z, z, 0.0, eps2, pos[lo:hi], vel[lo:hi])
        return r["seconds"]
    t0 = time.perf_counter()
    O.force(pos[lo:hi], vel[lo:hi], mass, pos, vel, eps2)
    return time.perf_counter() - t0


def cpu_rate(mass, pos, vel, eps2, ni_total, procs):
    """interactions/s of the reference CPU loop: ni_total sampled i x all j, split over `procs`
    processes (ph4 itself is single-threaded; its parallel mode is one MPI rank per core)."""
    import multiprocessing as mp
    from oracle import oracle as O
    kind = "reference" if O.ref_available() else "port"
    n = len(mass)
    per = max(1, ni_total // procs)
    jobs = [(kind, mass, pos, vel, eps2, k * per, (k + 1) * per) for k in range(procs)]
    if kind == "reference":
        O.ref()          # dlopen oracle/_ref/libph4ref.so in THIS process too, so the loaded-library record shows it
    if procs == 1:
        secs = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            secs = pool.map(_cpu_worker, jobs)
    dt = max(secs)
    return per * procs * float(n) / dt, kind, per * procs, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from amuse_b200 import plummer as P
    n = a.n
    mass, pos, vel = P.new_plummer_model(n, seed=a.seed)
    cores = os.cpu_count() or 1
    # size the sample so that one step is ~cpu_seconds/steps of work at ~5e7 interactions/s/core
    per_step_s = max(2.0, min(20.0, 60.0 / max(1, a.steps + a.warmup)))
    ni = int(max(cores, min(n, 5.0e7 * cores * per_step_s / n)))
    ni = max(cores, (ni // cores) * cores)
    for _ in range(a.warmup):
        cpu_rate(mass, pos, vel, a.eps2, max(cores, ni // 8), cores)
    rates = []
    t_tot = 0.0
    kind = "port"
    for _ in range(a.steps):
        r, kind, ni_used, dt = cpu_rate(mass, pos, vel, a.eps2, ni, cores)
        rates.append(r)
        t_tot += dt
    value = float(np.mean(rates))
    sample = "%d sampled i x %d j per step (%.1f%% of one N^2 sweep), %d processes" % (ni, n, 100.0 * ni / n, cores)
    line = {
        "impl": "reference", "metric": "interactions/s", "value": value, "unit": "interactions/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_tot / max(1, a.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "full i-block Hermite force sweep (acc, jerk, pot, nearest neighbour), "
                               "Plummer N=%d, eps2=%g; reference CPU loop on a bounded i-sample per step" % (n, a.eps2),
                   "n": n, "eps2": a.eps2, "sample_i_per_step": ni, "sample_fraction_of_sweep": ni / float(n),
                   "host_cores": cores},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for t, ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9 or not (t0 <= t <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _committed_traffic():
    """DRAM bytes of one force-kernel launch from the committed `ncu --set full` capture of this bench command
    (profiles/r02_force_kernel_ncu.txt: dram__bytes_read.sum + dram__bytes_write.sum, 16384 i x 1M j launch); DRAM
    counters cannot be read without the profiler, so the run itself reports the committed figure or null."""
    path = os.path.join(ROOT, "profiles", "r02_force_kernel_ncu.txt")
    try:
        tot = 0.0
        for l in open(path):
            if l.startswith("dram__bytes_read.sum") or l.startswith("dram__bytes_write.sum"):
                v, unit = l.split()[1], l.split()[2]
                tot += float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        if tot > 0:
            return tot, ("from the committed ncu --set full capture profiles/r02_force_kernel_ncu.txt (N=1M single-GPU launch "
                         "shape; algorithmic minimum ~ 50 MB j-set + boxes read once + per-split partials), not from this run")
    except Exception:
        pass
    return None, "not measured in the run (DRAM bytes need ncu)"


def _sample_errors(acc, jerk, pot, nn, ref, ids):
    """Per-particle relative errors of a sampled i-set against the oracle (north-star metric)."""
    ea = np.linalg.norm(acc - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ej = np.linalg.norm(jerk - ref["jerk"], axis=1) / np.linalg.norm(ref["jerk"], axis=1)
    ep = np.abs(pot - ref["pot"]) / np.abs(ref["pot"])
    return {"acc": float(ea.max()), "jerk": float(ej.max()), "pot": float(ep.max()),
            "nn_exact": float(np.mean(nn == ids[ref["nn"]]))}


def _extras(a, n_dev):
    """Driver-timed numbers of the other BASELINE configs (rank 0, N=1): block-step latency through the C ABI
    (oracle/g6_latency: a C caller, no Python in the loop) and the unmodified reference ph4 (oracle/_ref/libph4ref_gpu.so,
    its -DGPU objects linked to this library) at N=16k over a bounded interval, seconds per N-body time unit."""
    out = {}
    lib = os.path.join(ROOT, "amuse_b200", "csrc", "libsapporo.so")
    lat = os.path.join(ROOT, "oracle", "g6_latency")
    if os.path.exists(lat):
        table = {}
        for n in (16384, 131072):
            try:
                r = subprocess.run([lat, lib, str(n), "100"], capture_output=True, text=True, timeout=120)
                rows = {}
                for ln in r.stdout.splitlines():
                    f = ln.replace("|", " ").replace(":", " ").split()
                    if len(f) >= 4 and f[0] == "ni":
                        rows[f[1]] = {"force_call_us": float(f[2]), "with_j_updates_us": float(f[3])}
                table["N=%d" % n] = rows
            except Exception as e:  # harness missing or failed: report, do not fail the bench
                table["N=%d" % n] = repr(e)
        out["block_step_latency"] = {"what": "us per block step through the g6 C ABI (ni j-updates, set_ti, firsthalf, "
                                             "lasthalf2), C caller oracle/g6_latency.cc, uniform sphere, eps2=1e-4",
                                     "table": table}
    ref_gpu = os.path.join(ROOT, "oracle", "_ref", "libph4ref_gpu.so")
    if os.path.exists(ref_gpu):
        code = ("import sys, json; sys.path.insert(0, %r); from oracle import oracle as O; from amuse_b200 import plummer as P; "
                "m, x, v = P.new_plummer_model(16384, seed=1); "
                "r = O.ref_evolve(m, x, v, 0.0, 0.14, 0.0625, use_gpu=1, libname='libph4ref_gpu.so'); "
                "print('RESULT ' + json.dumps(r))") % ROOT
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            d = json.loads(line[-1][7:])
            out["ph4_s_per_unit"] = {"value": d["seconds"] / d["t"], "unit": "wall s per N-body time unit",
                                     "config": "unmodified ph4 (oracle/_ref/libph4ref_gpu.so) through the g6 ABI, Plummer "
                                               "N=16384, eps2=0, eta=0.14, t=0..%g" % d["t"],
                                     "block_steps": d["block_steps"], "particle_steps": d["particle_steps"],
                                     "dE_over_E": abs((d["E1"] - d["E0"]) / d["E0"])}
        except Exception as e:
            out["ph4_s_per_unit"] = {"value": None, "error": repr(e)}
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    from amuse_b200 import g6lib, plummer as P, sharding as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % a.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this library has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # ranks that only wait (while rank 0 drives all devices through the ABI) must wait on the CPU: an NCCL
        # barrier spins in a kernel on their GPU and would time-slice with the work rank 0 sends there
        cpu_group = dist.new_group(backend="gloo")

    n = a.n
    mass, pos, vel = P.new_plummer_model(n, seed=a.seed)      # identical on every rank
    ids = np.arange(1, n + 1, dtype=np.int32)
    if a.shuffle_ids:      # ids unrelated to the addresses (a caller that does not load in id order)
        ids = (np.random.RandomState(11).permutation(n) + 1).astype(np.int32)
    # Like ph4's MPI ranks (jdata.cc:56-67 + idata.cc:147-237) every rank holds ALL particles and owns a contiguous
    # j-domain -- here a window of the library's Morton-ordered j-memory (tile-aligned, define_window), so the close
    # pairs the library evaluates in FP64 live on one or two ranks and the neighbour bounds are global
    j0, j1 = 0, n
    os.environ.pop("G6_B200_DEVICES", None)
    g = g6lib.G6(local)
    L = g.L
    L.g6x_set_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), 1)
    L.g6x_set_j_offset(0)
    if a.variant:
        g.set_variant(a.variant)
    g.set_j_particles(ids, mass, pos, vel)
    njl = n
    w0, w1 = S.define_window(n, world, rank) if world > 1 else (0, 0)
    if world > 1:
        assert L.g6x_set_j_window(w0, w1) == 0
    nj_own = (w1 - w0) if world > 1 else n
    if world > 1 and a.exchange == "peer":
        S.attach_peers(L, n)      # CUDA IPC handles gathered over the process group
    npipes = g.npipes
    dchunk = L.g6x_device_chunk(n)           # i-particles per force-kernel launch on the device path
    n_launch = (n + dchunk - 1) // dchunk

    # device-resident i-block (all N particles; replicated on every rank like ph4's i-list)
    h_id = torch.from_numpy(ids).pin_memory()
    h_x = torch.from_numpy(pos).pin_memory()
    h_v = torch.from_numpy(vel).pin_memory()
    d_id, d_x, d_v = h_id.to(dev), h_x.to(dev), h_v.to(dev)
    d_sum = torch.empty((n, 7), dtype=torch.float64, device=dev)
    d_key = torch.empty(n, dtype=torch.int64, device=dev)
    d_nn = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    chunk_events = []

    def resolve(keys):
        L.g6x_resolve_nn(n, keys.data_ptr(), rank, d_nn.data_ptr())
        return d_nn

    def sweep(t, record=False, exchange=None):
        """predict + force sweep + cross-rank reduction; everything on the current stream."""
        exchange = exchange or a.exchange
        L.g6x_predict(njl, float(t))
        if record:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        # one call for the whole i-set: the library sorts it (Morton), cuts it into launches of `dchunk` i-particles
        # and combines the j-shards like idata.cc:284-313 on the device (sum, min key, id of the winner): either
        # fused into the force kernels (stores into the peers' exchange buffers over NVLink + one combine kernel) ...
        calc = L.g6x_calc_device_allreduce if (world > 1 and exchange == "peer") else L.g6x_calc_device
        calc(njl, n, d_id.data_ptr(), d_x.data_ptr(), d_v.data_ptr(), None, a.eps2, 1,
             d_sum.data_ptr(), d_key.data_ptr(), d_nn.data_ptr())
        if record:
            e1.record()
            chunk_events.append((e0, e1, n))
        if world > 1 and exchange == "nccl":      # ... or as three NCCL all-reduces
            S.combine_partials(d_sum, d_key, resolve)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a new time every step, so the predictor really runs each sweep; the particles carry no acc/jerk and the step is
    # 2^-40, so positions move by < 1e-12 and the parity sample below still compares against the t = 0 oracle
    tstep = 2.0 ** -40
    for w in range(a.warmup):
        flush.fill_(w)
        sweep(tstep * (w + 1))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = g.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(a.steps):
        flush.fill_(k)            # evict j/i data from L2 between timed iterations
        ev[k][0].record()
        sweep(tstep * (a.warmup + k + 1), record=True)
        ev[k][1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = g.launch_count() - launches0
    if world > 1 and a.exchange == "peer" and L.g6x_peer_error():
        raise SystemExit("bench.py: a combine kernel gave up waiting for a peer (g6x_peer_error)")
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_local = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    ms_t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    ms_per_step = ms_total / a.steps
    value = float(n) * float(n) / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (force_fast_kernel): per-launch algorithmic flop / mean launch duration
    # (events bracket the force launches of one sweep: Morton sort of the i-set, packing, neighbour-bound pre-pass,
    # force kernel, FP64 pair kernel and partial reduction; the predictor is outside them)
    kms = np.array([e0.elapsed_time(e1) for e0, e1, _ in chunk_events]) / n_launch
    flop_per_launch = FLOP_PER_INTERACTION * (float(n) / n_launch) * nj_own
    achieved = flop_per_launch / (kms.mean() * 1e-3) / 1e12
    kernel_share = float(kms.sum() * n_launch / ms_total)

    # ---- parity of what was just timed: a sampled i-set against the oracle (rank 0) ------------------------
    parity = None
    ref = None
    samp = np.sort(np.random.RandomState(5).choice(n, min(a.parity_sample, n), replace=False))
    if rank == 0 and a.parity_sample > 0:
        from oracle import oracle as O
        ref = O.force(pos[samp], vel[samp], mass, pos, vel, a.eps2, iid=ids[samp], jid=ids)
    torch.cuda.synchronize()
    if rank == 0 and ref is not None:
        sel = torch.from_numpy(samp).to(dev)
        s = d_sum[sel].cpu().numpy()
        parity = {"tol": 1e-6, "sample_i": int(len(samp)), "checker": "oracle/ (FP64 restatement of idata.cc:198-236)",
                  "device_path": _sample_errors(s[:, 0:3], s[:, 3:6], -s[:, 6], d_nn[sel].cpu().numpy(), ref, ids)}
    if world > 1 and a.exchange == "peer" and a.parity_sample > 0:
        # the same sweep with the exchange done by NCCL (three all-reduces): the fused exchange must agree
        keep = d_sum.clone()
        sweep(tstep * (a.warmup + a.steps + 1), exchange="nccl")
        torch.cuda.synchronize()
        if rank == 0:
            sel = torch.from_numpy(samp).to(dev)
            da = (keep[sel] - d_sum[sel]).abs().max(dim=0).values / d_sum[sel].abs().max(dim=0).values
            parity["fused_vs_nccl_exchange_max_rel_diff"] = float(da.max().item())
        del keep

    # measured FP32 FMA pipe peak (dependent-chain FFMA / FFMA2 microbenchmarks in the library)
    ffma = L.g6x_fp32_peak(0)
    ffma2 = L.g6x_fp32_peak(1)
    # predictor: HBM-bound kernel, 120 B read + 65 B written per j; L2 flushed between the timed launches
    pred_ms = L.g6x_time_predictor(njl, 20)
    pred_gbs = 185.0 * njl / (pred_ms * 1e-3) / 1e9 if pred_ms > 0 else None

    # j-update path (g6x_set_j_particles: host staging of 128 B records -> pinned batches -> H2D -> scatter_kernel):
    # reload this rank's whole j-shard and make it visible to the next force call
    torch.cuda.synchronize()
    tj0 = time.perf_counter()
    g.set_j_particles(ids[j0:j1], mass[j0:j1], pos[j0:j1], vel[j0:j1])
    L.g6x_predict(njl, 0.0)
    torch.cuda.synchronize()
    tj1 = time.perf_counter()
    j_update = {"value": njl / (tj1 - tj0), "unit": "particles/s", "bytes_per_particle": 128,
                "gbs": 128.0 * njl / (tj1 - tj0) / 1e9, "ms": 1e3 * (tj1 - tj0),
                "bound": "host staging loop + PCIe (8192-record batches uploaded while the rest is staged) + Morton "
                         "re-ordering of the j-memory (a reload of everything makes the order stale)"}

    # ---- end-to-end through the reference-facing API: the g6 C ABI with HOST arrays -------------------------
    # N = 1: this process' library instance.  N > 1: ONE process (rank 0) opens all N devices behind the same ABI
    # (G6_B200_DEVICES: j dealt out over the devices, partials gathered over peer memory) -- what a ph4 worker that
    # links the library gets; the other ranks release their devices and wait.
    e2e = None
    if not a.no_e2e:
        e2e_steps = max(1, min(a.steps, 2))
        g.close()
        barrier()
        torch.cuda.synchronize()
        how = None
        if rank == 0:
            if world > 1:
                os.environ["G6_B200_DEVICES"] = str(world)
            g2 = g6lib.G6(0)
            g2.set_j_particles(ids, mass, pos, vel)
            g2.set_ti(0.0)
            g2.calc(ids[:npipes], pos[:npipes], vel[:npipes], a.eps2)      # warm the ABI path (and order the j-memory)
            g2.synchronize()
            t0 = time.perf_counter()
            for k in range(e2e_steps):
                g2.set_ti(tstep * (100 + k))
                out = g2.calc(ids, pos, vel, a.eps2)
            dt = (time.perf_counter() - t0) / e2e_steps
            if parity is not None:
                parity["abi_e2e"] = _sample_errors(out["acc"][samp], out["jerk"][samp], out["pot"][samp], out["nn"][samp],
                                                   ref, ids)
            h2d, d2h = 64 * n * world, 60 * n
            how = "g6 C ABI (g6_set_ti_, g6calc_firsthalf_/g6calc_lasthalf2_), host double arrays, %d-particle chunks" % npipes
            if world > 1:
                how += ", ONE process driving %d devices (G6_B200_DEVICES=%d)" % (world, world)
            e2e = {"value": float(n) * float(n) / dt, "unit": "interactions/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "how": how}
            g2.close()
            os.environ.pop("G6_B200_DEVICES", None)
        if world > 1:
            dist.barrier(group=cpu_group)
    else:
        g.close()

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            ni_cpu = int(max(8, min(n, 5.0e7 * a.cpu_seconds / n)))
            r, kind, ni_used, dtc = cpu_rate(mass, pos, vel, a.eps2, ni_cpu, 1)
            cpu = {"value": r, "unit": "interactions/s", "cores": 1, "kind": kind,
                   "sample": "%d sampled i x %d j (%.2f%% of one sweep), %.1f s, ph4 force loop idata.cc:147-237" % (
                       ni_used, n, 100.0 * ni_used / n, dtc)}
        except Exception as e:  # the checker is optional infrastructure; never the measured path
            cpu = {"value": None, "unit": "interactions/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    extras = _extras(a, world) if (rank == 0 and world == 1 and a.extras) else {}

    ok = True
    if rank == 0:
        if parity is not None:
            for leg in ("device_path", "abi_e2e"):
                if leg in parity:
                    e = parity[leg]
                    parity[leg]["ok"] = bool(e["acc"] <= 1e-6 and e["jerk"] <= 1e-6 and e["pot"] <= 1e-6 and e["nn_exact"] >= 0.999)
                    ok = ok and parity[leg]["ok"]
            if "fused_vs_nccl_exchange_max_rel_diff" in parity:
                ok = ok and parity["fused_vs_nccl_exchange_max_rel_diff"] < 1e-7
            parity["ok"] = bool(ok)
        peak_src = "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json holds no FP32 figure); " \
                   "measured FFMA microbenchmark alongside"
        traffic, traffic_note = _committed_traffic()
        line = {
            "metric": "interactions/s", "value": value, "unit": "interactions/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 (double-single positions, f64 reduction, f64 close pairs)",
            "data": "synthetic",
            "config": {"workload": "full i-block Hermite force sweep (acc, jerk, pot, nearest neighbour), "
                                   "Plummer N=%d, eps2=%g, %d i-particles per launch, j sharded over %d GPU(s)" % (
                                       n, a.eps2, dchunk, world),
                       "n": n, "eps2": a.eps2, "npipes": npipes, "i_per_launch": dchunk, "l2": "256 MiB buffer written between timed steps",
                       "ids": "shuffled against the addresses" if a.shuffle_ids else "1..N in address order",
                       "j_domains": ("every rank holds all j and sums over its window of the Morton-ordered j-memory "
                                     "(%d slots here)" % nj_own) if world > 1 else "all j on one device",
                       "parallelism": "j-shard x%d + %s" % (world, "none" if world == 1 else (
                           "peer-memory exchange fused into the force kernels (NVLink stores + combine kernel)"
                           if a.exchange == "peer" else "3 NCCL all-reduces"))},
            "tflops_60": value * FLOP_PER_INTERACTION / 1e12,
            "frac_fp32_peak_nominal": value * FLOP_PER_INTERACTION / 1e12 / (NOMINAL_FP32_TFLOPS * world),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": NOMINAL_FP32_TFLOPS, "unit": "TFLOP/s",
                         "frac": achieved / NOMINAL_FP32_TFLOPS, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_note": traffic_note,
                         "peak_source": peak_src,
                         "kernel": "force_fast_kernel", "flop_per_launch": flop_per_launch,
                         "ms_per_launch": float(kms.mean()), "launches_timed": int(len(kms) * n_launch),
                         "kernel_share_of_step": kernel_share,
                         "measured_ffma_tflops": ffma, "measured_ffma2_tflops": ffma2,
                         "frac_of_measured_ffma": achieved / max(ffma, ffma2, 1e-9)},
            "predictor": {"bound": "hbm", "achieved": pred_gbs, "unit": "GB/s", "bytes_per_j": 185,
                          "ms_per_launch": pred_ms},
            "j_update": j_update,
            "e2e": e2e, "cpu_baseline": cpu, "parity": parity, "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(extras)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            if pred_gbs:
                line["predictor"]["peak"] = peaks.get("hbm_gbs")
                line["predictor"]["frac"] = pred_gbs / peaks.get("hbm_gbs")
        except Exception:
            pass
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: parity check failed (see the 'parity' block of the line above)")


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
