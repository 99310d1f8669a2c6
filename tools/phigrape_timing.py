"""BASELINE configs[2]: phiGRAPE's block-timestep loop (oracle/phigrape_replay.cc, the C++ restatement of the
Fortran caller) on a Plummer sphere through the g6 ABI: wall seconds per N-body unit, library share, and the
latency of one force call (set_ti + firsthalf + lasthalf2) by i-block size.
Usage: python tools/phigrape_timing.py N t_end [eps2] [max_block_steps] [b200|oracle|g6ref]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from amuse_b200 import plummer as P  # noqa: E402

n = int(sys.argv[1]); t_end = float(sys.argv[2])
eps2 = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
max_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
which = sys.argv[5] if len(sys.argv) > 5 else "b200"
lib = {"b200": "amuse_b200/csrc/libsapporo.so", "oracle": "oracle/liboracle_g6abi.so",
       "g6ref": "oracle/_ref/libg6ref.so"}[which]
m, x, v = P.new_plummer_model(n, seed=1)
inp = "/tmp/phigrape_%d.bin" % n
with open(inp, "wb") as f:
    np.int32(n).tofile(f); m.astype(np.float64).tofile(f)
    np.ascontiguousarray(x, dtype=np.float64).tofile(f); np.ascontiguousarray(v, dtype=np.float64).tofile(f)
out = subprocess.run([os.path.join(ROOT, "oracle", "phigrape_replay"), os.path.join(ROOT, lib), inp, repr(t_end),
                      repr(eps2), "0.02", "0.01", str(max_steps)], capture_output=True, text=True)
if out.returncode != 0:
    print("FAILED", out.stderr[-1000:]); sys.exit(1)
r = json.loads(out.stdout.strip().splitlines()[-1])
print("phiGRAPE loop N=%d eps2=%g on %s: %.3f s for %.5g time units = %.2f s per N-body unit; block steps %d, "
      "mean i-block %.1f, library force %.3f s + j-update %.3f s (host loop %.3f s), dE/E %.2e" % (
          n, eps2, which, r["seconds"], r["t"], r["seconds"] / max(r["t"], 1e-300), r["block_steps"],
          r["particle_steps"] / max(1, r["block_steps"]), r["lib_force_s"], r["lib_update_s"],
          r["seconds"] - r["lib_force_s"] - r["lib_update_s"], abs((r["E1"] - r["E0"]) / r["E0"])))
print("  force-call latency by i-block size (ni <= key: calls, mean us): " + json.dumps(r["latency_us"]))
