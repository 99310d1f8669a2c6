"""Quick force-kernel timing for tuning experiments (one launch shape, CUDA events, median of reps).
Usage: G6_B200_LIB=path/to/lib.so python tools/quick_force_bench.py [--n 262144] [--ni 16384] [--variants 6,7] [--refine 0,1]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=262144)
ap.add_argument("--ni", default="16384", help="comma-separated i-block sizes")
ap.add_argument("--variants", default="6,7")
ap.add_argument("--refine", default="1,0")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--nn", type=int, default=1)
ap.add_argument("--accuracy", type=int, default=1)
a = ap.parse_args()

dev = torch.device("cuda:0")
m, x, v = P.new_plummer_model(a.n, seed=1)
ids = np.arange(1, a.n + 1, dtype=np.int32)
g = g6lib.G6(0)
L = g.L
L.g6x_set_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), 1)
g.set_j_particles(ids, m, x, v)
L.g6x_predict(a.n, 0.0)
nis = [int(t) for t in a.ni.split(",")]
ni = max(nis)
d_id = torch.from_numpy(ids[:ni].copy()).to(dev)
d_x = torch.from_numpy(x[:ni].copy()).to(dev)
d_v = torch.from_numpy(v[:ni].copy()).to(dev)
d_sum = torch.empty((ni, 7), dtype=torch.float64, device=dev)
d_key = torch.empty(ni, dtype=torch.int64, device=dev)
d_nn = torch.empty(ni, dtype=torch.int32, device=dev)
ref = None
if a.accuracy:
    from oracle import oracle as O
    k = 128
    ref = O.force(x[:k], v[:k], m, x, v, 0.0, iid=ids[:k], jid=ids)
tag = os.path.basename(os.environ.get("G6_B200_LIB", "libsapporo.so"))
for var, ni in [(int(t), n_) for t in a.variants.split(",") for n_ in nis]:
    for rf in [int(t) for t in a.refine.split(",")]:
        g.set_variant(var)
        L.g6x_set_refine(rf)
        ts = []
        for r in range(a.reps + 2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            L.g6x_calc_device(a.n, ni, d_id.data_ptr(), d_x.data_ptr(), d_v.data_ptr(), None, 0.0, a.nn,
                              d_sum.data_ptr(), d_key.data_ptr(), d_nn.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        rate = ni * float(a.n) / (ms * 1e-3)
        acc = ""
        if ref is not None:
            s = d_sum[:128].cpu().numpy()
            ea = (np.linalg.norm(s[:, 0:3] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)).max()
            ej = (np.linalg.norm(s[:, 3:6] - ref["jerk"], axis=1) / np.linalg.norm(ref["jerk"], axis=1)).max()
            acc = " acc %.2e jerk %.2e" % (ea, ej)
        print("%-18s ni %6d variant %d refine %d nn %d: %.3f ms  %.4g int/s  %.1f%% of nominal%s" % (
            tag, ni, var, rf, a.nn, ms, rate, 100 * rate * 60 / 74.45e12, acc), flush=True)
g.close()
