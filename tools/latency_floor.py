"""Floor of the latency path on this box: k dependent empty kernels, the last raises a flag in mapped host memory,
the host spins on it (g6x_latency_probe).  Usage: python tools/latency_floor.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib  # noqa: E402

g = g6lib.G6(0)
for k in (1, 2, 3, 4):
    print("%d dependent empty launch(es) + host flag: %.2f us per round trip" % (k, g.L.g6x_latency_probe(k, 2000)))
g.close()
