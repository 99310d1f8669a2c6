"""Accuracy / speed of the FP64-radius cap (G6_B200_GMAX: group boxes a particle's FP64 radius may touch):
full-set maximum errors of N=20000 Plummer models (several seeds) through the ABI and the time of a sweep.
Usage: python tools/gmax_sweep.py [gmax list] [N] [seeds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

gl = [int(t) for t in (sys.argv[1] if len(sys.argv) > 1 else "96,192,1000000").split(",")]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
seeds = [int(t) for t in (sys.argv[3] if len(sys.argv) > 3 else "1,2,3,4").split(",")]
for seed in seeds:
    m, x, v = P.new_plummer_model(n, seed=seed)
    ids = np.arange(1, n + 1, dtype=np.int32)
    ref = O.force(x, v, m, x, v, 0.0, iid=ids, jid=ids)
    for gmax in gl:
        os.environ["G6_B200_GMAX"] = str(gmax)
        g = g6lib.G6(0)
        g.set_j_particles(ids, m, x, v)
        g.set_ti(0.0)
        g.calc(ids, x, v, 0.0)
        t0 = time.perf_counter()
        out = g.calc(ids, x, v, 0.0)
        dt = time.perf_counter() - t0
        g.close()
        ea = np.linalg.norm(out["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
        ej = np.linalg.norm(out["jerk"] - ref["jerk"], axis=1) / np.linalg.norm(ref["jerk"], axis=1)
        print("N=%d seed %d gmax %7d: sweep %.2f ms | acc max %.2e jerk max %.2e, %d particles above 5e-7, nn exact %s" % (
            n, seed, gmax, 1e3 * dt, ea.max(), ej.max(), int((ej > 5e-7).sum()), bool(np.array_equal(out["nn"], ref["nn"]))), flush=True)
