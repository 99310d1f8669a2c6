"""BASELINE configs[3] through the product API: ONE process, the g6 C ABI, G6_B200_DEVICES devices.
Plummer N (default 1M) + 10 % primordial binaries (test_multiples2.py:236-304), eps2 = 0.

  (i)   full sweep of all particles through g6calc_firsthalf_/lasthalf2_ with host arrays: interactions/s, parity of a
        sampled i-subset (binary members included) against the FP64 oracle, Newton's third law over the whole sweep,
        binaries found as mutual nearest neighbours;
  (ii)  neighbour lists (h2 = min(8 dnn^2, 1), gpu.cc:629-630) of a sampled i-block against the oracle;
  (iii) the UNMODIFIED ph4 integrator (oracle/_ref/libph4ref_gpu.so: its -DGPU objects linked to this library) for a
        bounded number of block steps, with its own close-encounter management on (manage_encounters = 1, what the
        standalone driver defaults to, parallel_hermite_4.cc:306) and off: seconds per block step and, by
        extrapolation, per N-body time unit.

Usage: G6_B200_DEVICES=8 python tools/config4_abi.py [N] [ph4 block steps]      prints one 'CONFIG4-ABI {json}' line"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from amuse_b200 import g6lib, plummer as P  # noqa: E402
from oracle import oracle as O  # noqa: E402
from helpers import check_forces, check_nn, error_report  # noqa: E402


def main():
    n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    ph4_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    m, x, v = P.new_plummer_model(n0, seed=1)
    ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1)
    n = len(m)
    g = g6lib.G6(0)
    ndev = g.L.g6x_device_count_open()
    res = {"config": "Plummer N=%d + 10%% binaries = %d particles, eps2=0, ONE process, g6 C ABI, %d device(s)" % (n0, n, ndev),
           "max_id": int(ids.max())}
    t0 = time.perf_counter()
    g.set_j_particles(ids, m, x, v)
    g.set_ti(0.0)
    g.calc(ids[:g.npipes], x[:g.npipes], v[:g.npipes], 0.0)
    g.synchronize()
    res["load_and_first_chunk_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = g.calc(ids, x, v, 0.0)
    dt = time.perf_counter() - t0
    res["sweep_s"] = dt
    res["interactions_per_s"] = float(n) * n / dt
    acc, jerk, nn = out["acc"], out["jerk"], out["nn"]
    res["newton3_acc"] = float(np.abs((m[:, None] * acc).sum(axis=0)).max() / (m * np.linalg.norm(acc, axis=1)).sum())
    res["newton3_jerk"] = float(np.abs((m[:, None] * jerk).sum(axis=0)).max() / (m * np.linalg.norm(jerk, axis=1)).sum())
    nbin = n - n0
    prim = np.arange(0, n0, n0 // nbin)[:nbin]
    res["binaries_mutual_nn"] = float(np.mean((nn[prim] == ids[n0:]) & (nn[n0:] == ids[prim])))
    rnd = np.random.RandomState(0)
    samp = np.sort(np.concatenate([rnd.choice(n0, 480, replace=False), prim[:16], n0 + np.arange(16)]))
    ref = O.force(x[samp], v[samp], m, x, v, 0.0, iid=ids[samp], jid=ids, scales=True)
    got = dict(acc=acc[samp], jerk=jerk[samp], pot=out["pot"][samp])
    res["parity_sample"] = error_report(got, ref)
    try:
        check_forces(got, ref, what="config 4 sample")
        check_nn(nn[samp], ref["nn"], ids, x[samp], x)
        res["parity_ok"] = True
    except AssertionError as e:
        res["parity_ok"] = False
        res["parity_msg"] = str(e)
    # neighbour lists of a sampled block
    sub = samp[:128]
    h2 = np.minimum(8 * ref["dnn"][:128] ** 2, 1.0)
    t0 = time.perf_counter()
    g.calc(ids[sub], x[sub], v[sub], 0.0, h2=h2)
    res["ngb_overflow"] = int(g.read_neighbour_list())
    bad, lens = 0, []
    for k, i in enumerate(sub):
        rc, c, l = g.get_neighbour_list(k)
        lens.append(c)
        cr, lr = O.neighbours(int(ids[i]), x[i], h2[k], ids, m, x)
        r2 = ((x - x[i]) ** 2).sum(axis=1)
        edge = set(ids[np.abs(r2 - h2[k]) <= 1e-6 * h2[k]].tolist())
        if not (set(l.tolist()) ^ set(lr.tolist()) <= edge):
            bad += 1
    res["ngb_seconds"] = time.perf_counter() - t0
    res["ngb_mean_len"] = float(np.mean(lens))
    res["ngb_lists_wrong"] = bad
    g.close()
    # the unmodified ph4 on this library (own process each: ph4 keeps function-static state)
    if ph4_steps > 0 and O.ref_available("libph4ref_gpu.so"):
        code = ("import sys, json, numpy as np; sys.path.insert(0, %r); from oracle import oracle as O; "
                "from amuse_b200 import plummer as P; m, x, v = P.new_plummer_model(%d, seed=1); "
                "ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1); "
                "r = O.ref_evolve_enc(m, x, v, 0.0, 0.14, 1.0, manage_encounters=%%d, ids=ids, use_gpu=True, "
                "max_block_steps=%d); print('RESULT ' + json.dumps({k: float(v_) for k, v_ in r.items()}))") % (ROOT, n0, ph4_steps)
        res["ph4"] = {}
        for mode in ((1,) if os.environ.get("CONFIG4_ENC_ONLY") else (1, 0)):
            t0 = time.perf_counter()
            p = subprocess.run([sys.executable, "-c", code % mode], capture_output=True, text=True, timeout=3000)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if not line:
                res["ph4"]["manage_encounters=%d" % mode] = {"error": (p.stderr or p.stdout)[-400:]}
                continue
            r = json.loads(line[-1][7:])
            res["ph4"]["manage_encounters=%d" % mode] = {
                "block_steps": r["block_steps"], "particle_steps": r["particle_steps"], "t_reached": r["t"],
                "seconds_in_advance_loop": r["seconds"], "us_per_block_step": 1e6 * r["seconds"] / max(1.0, r["block_steps"]),
                "s_per_nbody_unit_extrapolated": r["seconds"] / r["t"] if r["t"] > 0 else None,
                "particles_left": r["nj_left"], "dE_over_E": abs((r["E1"] - r["E0"]) / r["E0"]),
                "wall_s_total_incl_setup_sweeps": time.perf_counter() - t0}
        res["ph4"]["note"] = ("standalone ph4 (ref_driver.cc around the unmodified jdata/idata/scheduler objects), eps2 = 0; "
                              "manage_encounters = 1 is ph4's own pairwise close-encounter treatment "
                              "(close_encounter.cc:66-73), 0 switches it off")
    print("CONFIG4-ABI " + json.dumps(res))


if __name__ == "__main__":
    main()
