"""The reference ph4 integrator driving the B200 library through the g6 ABI (BASELINE configs[0]/[1]
shape): the UNMODIFIED ph4 sources compiled -DGPU -DNOMPI (oracle/_ref/libph4ref_gpu.so, linked
against amuse_b200/csrc/libsapporo.so) versus the same sources in CPU mode (oracle/_ref/libph4ref.so).
Each evolve runs in its own process: ph4 keeps function-static state (gpu.cc:302-330)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys, json
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import oracle as O
from amuse_b200 import plummer as P
m, x, v = P.new_plummer_model(%(n)d, seed=1, do_scale=%(scale)s)
r = O.ref_evolve(m, x, v, %(eps2)g, 0.14, %(t)g, use_gpu=%(gpu)d, libname=%(lib)r)
print("RESULT " + json.dumps(r))
"""


def _run(n, eps2, t, gpu, scale=False):
    lib = "libph4ref_gpu.so" if gpu else "libph4ref.so"
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", lib)):
        pytest.skip("oracle/_ref/%s not built (make -C oracle ref refgpu in the dev container)" % lib)
    code = CODE % dict(root=ROOT, n=n, eps2=eps2, t=t, gpu=int(gpu), lib=lib, scale=scale)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


@pytest.mark.parametrize("n,eps2,t", [(1024, 1e-4, 0.25), (1024, 0.0, 0.125)])
def test_ph4_gpu_mode_tracks_cpu_mode(n, eps2, t):
    cpu = _run(n, eps2, t, gpu=False, scale=True)
    gpu = _run(n, eps2, t, gpu=True, scale=True)
    print("ph4 N=%d eps2=%g t=%g: CPU E0=%.12f E1=%.12f steps=%d/%d %.2fs | g6-B200 E0=%.12f E1=%.12f steps=%d/%d %.2fs"
          % (n, eps2, t, cpu["E0"], cpu["E1"], cpu["block_steps"], cpu["particle_steps"], cpu["seconds"],
             gpu["E0"], gpu["E1"], gpu["block_steps"], gpu["particle_steps"], gpu["seconds"]))
    assert abs(gpu["E0"] - cpu["E0"]) < 2e-7 * abs(cpu["E0"])            # same initial energy (FP32-level forces)
    assert abs(gpu["E1"] - gpu["E0"]) < 2e-5 * abs(gpu["E0"])            # energy conserved like the CPU run
    assert abs(cpu["E1"] - cpu["E0"]) < 2e-5 * abs(cpu["E0"])
    # chaotic divergence allows step counts to differ slightly, not grossly
    assert abs(gpu["particle_steps"] - cpu["particle_steps"]) < 0.05 * cpu["particle_steps"]
    assert gpu["t"] >= t
