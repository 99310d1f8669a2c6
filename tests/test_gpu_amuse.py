"""ph4 end-to-end through the AMUSE framework: the unmodified Python coupling layer
(amuse.rfi sockets channel, amuse_ph4/interface.py) starts `ph4_sapporo_worker` -- the reference's generated worker +
interface.cc + ph4 -DGPU objects, linked against amuse_b200/csrc/libsapporo.so -- and runs the reference's own
GPU-vs-CPU test (src/amuse_ph4/tests/test_ph4.py:848-872) plus an evolve.  Everything AMUSE-side comes from the
git-ignored oracle/_ref/amuse (built by `make -C oracle amuse` in the dev container; docutils is stubbed by
tests/amuse_stub because this image lacks it)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AMUSE = os.path.join(ROOT, "oracle", "_ref", "amuse")


def test_ph4_sapporo_worker_through_amuse_matches_cpu_worker():
    if not os.path.exists(os.path.join(AMUSE, "ph4_sapporo_worker")):
        pytest.skip("oracle/_ref/amuse not built (make -C oracle amuse in the dev container)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "amuse_stub"), os.path.join(AMUSE, "py")])
    env.pop("G6_B200_DEVICES", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "amuse_ph4_script.py"), AMUSE],
                         capture_output=True, text=True, timeout=900, env=env)
    lines = {l.split()[1]: json.loads(l.split(" ", 2)[2]) for l in out.stdout.splitlines() if l.startswith("AMUSE-PH4 ")}
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("AMUSE-PH4 ")))
    assert out.returncode == 0 and "test22_gpu" in lines and "evolve" in lines, out.stdout[-3000:] + out.stderr[-3000:]
    # the reference asserts 1e-5 (assertAlmostRelativeEquals(..., 5)); this library is held to its own 1e-6
    assert lines["test22_gpu"]["max_rel_diff_potential"] < 1e-6
    ev = lines["evolve"]
    assert abs(ev["gpu N=1024"]["E0"] - ev["cpu N=1024"]["E0"]) < 1e-6 * abs(ev["cpu N=1024"]["E0"])
    for k, r in ev.items():
        assert r["dE_over_E"] < 2e-5, (k, r)
