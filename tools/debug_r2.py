"""Bisects a force mismatch: runs one case through the ABI with the pair classification switched off piece by piece
and prints the worst particles with the pair contribution that would explain their error."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P
from oracle import oracle as O

case = sys.argv[1] if len(sys.argv) > 1 else "binaries"
if case == "binaries":
    m, x, v = P.new_plummer_model(8000, seed=2)
    ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1)
    eps2 = 0.0
else:
    n = int(case)
    m, x, v = P.new_plummer_model(n, seed=8)
    ids = np.arange(1, n + 1, dtype=np.int32)
    eps2 = 1e-4
n = len(m)
ref = O.force(x, v, m, x, v, eps2)
g = g6lib.G6(0)
for label, K, farc, var in [("default", 16, 0.125, 0), ("K=0", 0, 0.125, 0), ("no FAR", 16, 1e9, 0), ("K=0 no FAR", 0, 1e9, 0),
                            ("masked P2", 16, 0.125, 7)]:
    g.nj = 0
    g.set_variant(var)
    g.set_close_factor(K, farc)
    g.set_j_particles(ids, m, x, v)
    g.set_ti(0.0)
    out = g.calc(ids, x, v, eps2)
    ea = np.linalg.norm(out["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ej = np.linalg.norm(out["jerk"] - ref["jerk"], axis=1) / np.linalg.norm(ref["jerk"], axis=1)
    ep = np.abs(out["pot"] - ref["pot"]) / np.abs(ref["pot"])
    nbad = (ea > 1e-5).sum()
    print("%-12s acc max %.2e jerk max %.2e pot max %.2e; particles with acc err > 1e-5: %d; nn exact %.4f" % (
        label, ea.max(), ej.max(), ep.max(), nbad, (out["nn"] == ids[ref["nn"]]).mean()), flush=True)
    for i in np.argsort(-ea)[:min(6, nbad)]:
        d = out["acc"][i] - ref["acc"][i]
        X = x - x[i]; r2 = (X * X).sum(1); r2[i] = np.inf
        contrib = (m / r2 ** 1.5)[:, None] * X
        # which single pair explains the difference (missing: d = -contrib, doubled: d = +contrib)?
        miss = np.linalg.norm(contrib + d, axis=1); dbl = np.linalg.norm(contrib - d, axis=1)
        jm, jd = miss.argmin(), dbl.argmin()
        print("   i=%d id=%d err %.2e |a_ref| %.3e |d| %.3e nn(ref) id %d r %.3e | best 'missing pair' j=%d (id %d, r %.2e) resid %.2e | "
              "best 'doubled pair' j=%d resid %.2e | dpot %.3e" % (
                  i, ids[i], ea[i], np.linalg.norm(ref["acc"][i]), np.linalg.norm(d), ids[ref["nn"][i]],
                  np.sqrt(r2[ref["nn"][i]]), jm, ids[jm], np.sqrt(r2[jm]), miss[jm] / np.linalg.norm(d), jd,
                  dbl[jd] / np.linalg.norm(d), out["pot"][i] - ref["pot"][i]))
g.close()
