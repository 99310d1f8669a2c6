import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
from amuse_b200 import g6lib, plummer as P
from oracle import oracle as O
m, x, v = P.new_plummer_model(8000, seed=2)
ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1)
ref = O.force(x, v, m, x, v, 0.0)
m2, x2, v2 = P.new_plummer_model(20000, seed=4); ids2 = np.arange(1, 20001, dtype=np.int32)
ref2 = O.force(x2, v2, m2, x2, v2, 0.0)
g = g6lib.G6(0)
for K in (32.0, 0.0):
  for farc in (0.125, 0.25, 0.5):
    for name, (I, M_, X, V, R) in {"binaries 8.8k": (ids, m, x, v, ref), "plummer 20k": (ids2, m2, x2, v2, ref2)}.items():
        g.nj = 0; g.set_close_factor(K, farc); g.set_j_particles(I, M_, X, V); g.set_ti(0.0)
        out = g.calc(I, X, V, 0.0)
        ea = np.linalg.norm(out["acc"] - R["acc"], axis=1) / np.linalg.norm(R["acc"], axis=1)
        ej = np.linalg.norm(out["jerk"] - R["jerk"], axis=1) / np.linalg.norm(R["jerk"], axis=1)
        ep = np.abs(out["pot"] - R["pot"]) / np.abs(R["pot"])
        print("K=%g farc=%g %s: acc %.2e jerk %.2e pot %.2e" % (K, farc, name, ea.max(), ej.max(), ep.max()), flush=True)
g.close()
