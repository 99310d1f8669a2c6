"""Wall time per g6calc_firsthalf_/lasthalf2_ chunk through the ABI at N = 1M (or --n), for the devices
G6_B200_DEVICES asks for; prints the library's host-phase trace with G6_B200_TRACE=1."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
m, x, v = P.new_plummer_model(n, seed=1)
ids = np.arange(1, n + 1, dtype=np.int32)
g = g6lib.G6(0)
t0 = time.perf_counter(); g.set_j_particles(ids, m, x, v); g.set_ti(0.0)
g.calc(ids[:g.npipes], x[:g.npipes], v[:g.npipes], 0.0); g.synchronize()
print("devices %d: load + first chunk %.2f s" % (g.L.g6x_device_count_open(), time.perf_counter() - t0), flush=True)
ni = nch * g.npipes
t0 = time.perf_counter(); out = g.calc(ids[:ni], x[:ni], v[:ni], 0.0); dt = time.perf_counter() - t0
print("devices %d: %.2f ms per chunk of %d (%.4g interactions/s)" % (g.L.g6x_device_count_open(), 1e3 * dt / nch, g.npipes, ni * float(n) / dt), flush=True)
g.close()
