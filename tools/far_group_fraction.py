"""Design probe for DESIGN.md section 8 (no GPU): if the j-memory and the i-blocks were kept in Morton order, what
fraction of the speculative kernel's (warp of 64 i) x (group of 32 j) blocks would have NO pair closer than r_far --
i.e. could be evaluated with single-precision position differences (hi parts only, 6 of 37 operations saved)?
Compared with the caller's (random) order the library sees today.  Usage: python tools/far_group_fraction.py [N] [warps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import plummer as P  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
nwarps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
m, x, v = P.new_plummer_model(n, seed=1)


def morton(x, bits=16):
    lo, hi = x.min(axis=0), x.max(axis=0)
    q = np.minimum(((x - lo) / (hi - lo) * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    key = np.zeros(len(x), dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return key


rnd = np.random.RandomState(0)
for name, order in (("caller (random) order", np.arange(n)), ("Morton order", np.argsort(morton(x)))):
    xs = x[order]
    ng = n // 32
    groups = xs[:ng * 32].reshape(ng, 32, 3)
    gc = groups.mean(axis=1)
    gr = np.sqrt(((groups - gc[:, None, :]) ** 2).sum(axis=2)).max(axis=1)      # bounding sphere of each j-group
    warps = rnd.choice(n // 64, nwarps, replace=False)
    res = {0.02: [], 0.05: [], 0.1: []}
    for w in warps:
        xi = xs[w * 64:(w + 1) * 64]
        wc = xi.mean(axis=0)
        wr = np.sqrt(((xi - wc) ** 2).sum(axis=1)).max()
        gap = np.sqrt(((gc - wc) ** 2).sum(axis=1)) - gr - wr                      # lower bound of the closest pair
        # exact minimum only where the bound is inconclusive
        for rf in res:
            far = gap > rf
            idx = np.nonzero(~far & (gap > -1e9))[0]
            if len(idx) > 4000:
                idx = rnd.choice(idx, 4000, replace=False); scale = (~far).sum() / 4000.0
            else:
                scale = 1.0
            d2 = ((groups[idx][:, None, :, :] - xi[None, :, None, :]) ** 2).sum(axis=3).min(axis=(1, 2))
            far_extra = (d2 > rf * rf).sum() * scale
            res[rf].append((far.sum() + far_extra) / ng)
    print("%-22s N=%d: fraction of (64 i) x (32 j) blocks with no pair closer than r_far: " % (name, n) +
          ", ".join("r_far=%g: %.1f%%" % (rf, 100 * np.mean(val)) for rf, val in res.items()))
