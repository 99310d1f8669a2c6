"""Round-2 force-kernel diagnostics: timing of one sweep on the device path (globally Morton-sorted i-set) and through
the ABI (chunks of g6_npipes() particles in caller order), with the fraction of (warp x group) blocks the kernel took
FAR / NEAR / CLOSE (needs the -DG6_STATS build: G6_B200_LIB=amuse_b200/csrc/libsapporo_stats.so) and the accuracy of a
sampled i-subset against the oracle, for a list of (K, FARC) settings.
Usage: python tools/block_stats.py [--n 1048576] [--k 16,8,0] [--farc 0.125] [--abi-chunks 4] [--order caller|morton]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--k", default="16")
ap.add_argument("--farc", default="0.125")
ap.add_argument("--eps2", type=float, default=0.0)
ap.add_argument("--abi-chunks", type=int, default=4)
ap.add_argument("--sample", type=int, default=2048)
ap.add_argument("--binaries", type=float, default=0.0)
ap.add_argument("--jshard", type=int, default=1, help="load only the first 1/jshard of the particles as j (one rank of a sharded run)")
a = ap.parse_args()

dev = torch.device("cuda:0")
m, x, v = P.new_plummer_model(a.n, seed=1)
ids = np.arange(1, a.n + 1, dtype=np.int32)
if a.binaries > 0:
    ids, m, x, v = P.add_binaries(m, x, v, a.binaries, seed=7)
n = len(m)
g = g6lib.G6(0)
L = g.L
L.g6x_set_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), 1)
nj = n // a.jshard
g.set_j_particles(ids[:nj], m[:nj], x[:nj], v[:nj])
d_id = torch.from_numpy(ids).to(dev)
d_x = torch.from_numpy(x).to(dev)
d_v = torch.from_numpy(v).to(dev)
d_sum = torch.empty((n, 7), dtype=torch.float64, device=dev)
d_key = torch.empty(n, dtype=torch.int64, device=dev)
d_nn = torch.empty(n, dtype=torch.int32, device=dev)
st = (C.c_ulonglong * 8)()
have_stats = L.g6x_block_stats(st) == 0
from oracle import oracle as O  # noqa: E402
rnd = np.random.RandomState(5)
samp = np.sort(rnd.choice(n, min(a.sample, n), replace=False))
ref = O.force(x[samp], v[samp], m[:nj], x[:nj], v[:nj], a.eps2, iid=ids[samp], jid=ids[:nj])


def errs(acc, jerk, pot):
    ea = np.linalg.norm(acc - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ej = np.linalg.norm(jerk - ref["jerk"], axis=1) / np.linalg.norm(ref["jerk"], axis=1)
    ep = np.abs(pot - ref["pot"]) / np.abs(ref["pot"])
    return "acc %.2e jerk %.2e (p99 %.2e) pot %.2e" % (ea.max(), ej.max(), np.percentile(ej, 99), ep.max())


def stats():
    if not have_stats:
        return ""
    L.g6x_block_stats(st)
    tot = float(sum(st[:3])) or 1.0
    return " | blocks FAR %.2f%% NEAR %.2f%% CLOSE %.3f%% (NEAR redone %.4f%%), FP64 pairs queued %d (%d appends met a full list)" % (
        100 * st[0] / tot, 100 * st[1] / tot, 100 * st[2] / tot, 100 * st[3] / tot, st[4], st[5])


tag = os.path.basename(os.environ.get("G6_B200_LIB", "libsapporo.so"))
print("%s N=%d eps2=%g npipes=%d" % (tag, n, a.eps2, g.npipes), flush=True)
for K in [float(t) for t in a.k.split(",")]:
    for farc in [float(t) for t in a.farc.split(",")]:
        g.set_close_factor(K, farc)
        ts = []
        for r in range(3):
            L.g6x_predict(nj, 0.0)
            torch.cuda.synchronize()
            if r == 2:
                stats()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            L.g6x_calc_device(nj, n, d_id.data_ptr(), d_x.data_ptr(), d_v.data_ptr(), None, a.eps2, 1,
                              d_sum.data_ptr(), d_key.data_ptr(), d_nn.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts[1:])
        s = d_sum[torch.from_numpy(samp).to(dev)].cpu().numpy()
        print("device path  K=%-4g farc=%-6g: %.1f ms/sweep %.4g int/s %.1f%% of nominal | %s%s" % (
            K, farc, ms, float(n) * nj / (ms * 1e-3), 100 * float(n) * nj * 60 / (ms * 1e-3) / 74.45e12,
            errs(s[:, 0:3], s[:, 3:6], -s[:, 6]), stats()), flush=True)
        if a.abi_chunks > 0:
            L.g6x_set_stream(None, 0)
            nch = min(a.abi_chunks, (n + g.npipes - 1) // g.npipes)
            ni = min(n, nch * g.npipes)
            g.set_ti(0.0)
            g.calc(ids[:g.npipes], x[:g.npipes], v[:g.npipes], a.eps2)
            stats()
            t0 = time.perf_counter()
            out = g.calc(ids[:ni], x[:ni], v[:ni], a.eps2)
            dt = time.perf_counter() - t0
            sel = samp[samp < ni]
            sub = np.searchsorted(samp, sel)
            ea = np.linalg.norm(out["acc"][sel] - ref["acc"][sub], axis=1) / np.linalg.norm(ref["acc"][sub], axis=1)
            ej = np.linalg.norm(out["jerk"][sel] - ref["jerk"][sub], axis=1) / np.linalg.norm(ref["jerk"][sub], axis=1)
            print("ABI, %d chunks of %d (caller order): %.2f ms/chunk %.4g int/s %.1f%% of nominal | acc %.2e jerk %.2e%s" % (
                nch, g.npipes, 1e3 * dt / nch, ni * float(n) / dt, 100 * ni * float(n) * 60 / dt / 74.45e12,
                ea.max() if len(sel) else 0, ej.max() if len(sel) else 0, stats()), flush=True)
            L.g6x_set_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), 1)
g.close()
