"""BASELINE configs[0]/[1]: the reference ph4 integrator (unmodified sources) in CPU mode and through the g6 ABI
on the B200 library: wall seconds per N-body time unit.  Usage: python tools/ph4_timing.py N t_end [cpu|gpu|both]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, json
sys.path.insert(0, %(root)r)
from oracle import oracle as O
from amuse_b200 import plummer as P
m, x, v = P.new_plummer_model(%(n)d, seed=1)
r = O.ref_evolve(m, x, v, %(eps2)g, 0.14, %(t)g, use_gpu=%(gpu)d, libname=%(lib)r)
print("RESULT " + json.dumps(r))
"""
n = int(sys.argv[1]); t = float(sys.argv[2]); which = sys.argv[3] if len(sys.argv) > 3 else "both"
eps2 = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
for gpu in ([0, 1] if which == "both" else [1 if which == "gpu" else 0]):
    lib = "libph4ref_gpu.so" if gpu else "libph4ref.so"
    out = subprocess.run([sys.executable, "-c", CODE % dict(root=ROOT, n=n, eps2=eps2, t=t, gpu=gpu, lib=lib)],
                         capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        print("FAILED", out.stderr[-500:]); continue
    r = json.loads(line[-1][7:])
    print("ph4 N=%d eps2=%g %s: %.3f s for %.4g time units = %.2f s per N-body unit; block steps %d, particle steps %d, "
          "mean i-block %.1f, dE/E %.2e" % (n, eps2, "g6-B200" if gpu else "CPU    ", r["seconds"], r["t"], r["seconds"] / r["t"],
                                            r["block_steps"], r["particle_steps"], r["particle_steps"] / max(1, r["block_steps"]),
                                            abs((r["E1"] - r["E0"]) / r["E0"])))
