"""One ABI chunk and one device-path chunk at N = 1M for K = 16 and K = 0, to be run under
ncu --metrics gpu__time_duration.sum (per-kernel durations of the round-2 pipeline)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m, x, v = P.new_plummer_model(n, seed=1)
ids = np.arange(1, n + 1, dtype=np.int32)
g = g6lib.G6(0); L = g.L
g.set_j_particles(ids, m, x, v)
dev = torch.device("cuda:0")
nd = 18944
order = np.argsort(x[:, 0] + 3 * x[:, 1] + 7 * x[:, 2])   # any fixed order; the library sorts
d_id = torch.from_numpy(ids[:nd].copy()).to(dev); d_x = torch.from_numpy(x[:nd].copy()).to(dev); d_v = torch.from_numpy(v[:nd].copy()).to(dev)
d_sum = torch.empty((nd, 7), dtype=torch.float64, device=dev); d_key = torch.empty(nd, dtype=torch.int64, device=dev); d_nn = torch.empty(nd, dtype=torch.int32, device=dev)
for K in (16.0, 0.0):
    g.set_close_factor(K, 0.125)
    g.set_ti(0.0)
    for rep in range(2):
        g.calc(ids[:16384], x[:16384], v[:16384], 0.0)
    torch.cuda.synchronize()
print("done")
g.close()
