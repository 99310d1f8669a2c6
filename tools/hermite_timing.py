"""Device-resident Hermite integrator (g6x_hermite_evolve) vs the unmodified ph4 integrator through the g6 ABI and
in CPU mode: wall seconds per N-body time unit on a Plummer sphere.  Usage: python tools/hermite_timing.py N t_end [eps2]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P  # noqa: E402

n = int(sys.argv[1]); t_end = float(sys.argv[2]); eps2 = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
m, x, v = P.new_plummer_model(n, seed=1)
ids = np.arange(1, n + 1, dtype=np.int32)
g = g6lib.G6(0)
g.set_j_particles(ids, m, x, v)
g.L.g6x_hermite_init(n, 0.0, 0.14, eps2, None)
st = np.zeros(4)
g.L.g6x_hermite_evolve(n, t_end, 0.14, eps2, 0, st)
print("device-resident Hermite N=%d eps2=%g: %.3f s for %.4g time units = %.3f s per N-body unit; block steps %d, "
      "particle steps %d, mean i-block %.1f, %.1f us per block step" % (
          n, eps2, st[3], st[0], st[3] / st[0], st[1], st[2], st[2] / st[1], 1e6 * st[3] / st[1]))
g.close()
