"""Block-timestep regime (BASELINE configs[2] shape): latency of one g6 force evaluation through the
C ABI (set_ti + ni x set_j_particle + firsthalf + lasthalf2) for small i-blocks, host arrays in and out.
Usage: python tools/latency_bench.py [--n 131072]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--reps", type=int, default=200)
a = ap.parse_args()
n = a.n
m, x, v = P.new_plummer_model(n, seed=1)
ids = np.arange(1, n + 1, dtype=np.int32)
g = g6lib.G6(0)
L = g.L
g.set_j_particles(ids, m, x, v)
g.set_ti(0.0)
g.calc(ids[:1024], x[:1024], v[:1024], 1e-4)
cid = g.cid
z3 = np.zeros((g.npipes, 3)); z1 = np.zeros(g.npipes)
print("N = %d" % n)
for ni in (1, 4, 16, 42, 128, 512, 2048, 8192):
    rnd = np.random.RandomState(ni)
    sel = np.sort(rnd.choice(n, ni, replace=False))
    idc, xc, vc = ids[sel].copy(), x[sel].copy(), v[sel].copy()
    acc = np.empty((ni, 3)); jerk = np.empty((ni, 3)); pot = np.empty(ni); inn = np.empty(ni, dtype=np.int32)
    h2 = np.zeros(ni)
    e = C.c_double(1e-4); cnj = C.c_int(n); cn = C.c_int(ni)
    tset = tforce = 0.0
    for r in range(a.reps + 5):
        ti = C.c_double(1e-6 * (r + 1))
        t0 = time.perf_counter()
        L.g6_set_ti_(C.byref(cid), C.byref(ti))
        L.g6calc_firsthalf_(C.byref(cid), C.byref(cnj), C.byref(cn), idc, xc, vc, z3[:ni], z3[:ni], z1[:ni], C.byref(e), h2)
        L.g6calc_lasthalf2_(C.byref(cid), C.byref(cnj), C.byref(cn), idc, xc, vc, C.byref(e), h2, acc, jerk, pot, inn)
        t1 = time.perf_counter()
        # the caller's j-update of the block it just advanced (idata::update_gpu, gpu.cc:163-230)
        L.g6x_set_j_particles(ni, idc.ctypes.data, 0, idc, None, m[sel].copy(), None, None, vc, xc) if False else None
        if r >= 5:
            tforce += t1 - t0
    us = 1e6 * tforce / a.reps
    print("ni %5d: %8.1f us per force call  (%.3g interactions/s)" % (ni, us, ni * float(n) / (us * 1e-6)))
g.close()
