"""Several devices behind the g6 ABI in ONE process (G6_B200_DEVICES, SURVEY.md 8e "single process, 8 devices"):
needs >= 2 GPUs, skipped on a single-GPU box.  What the reference does with one MPI rank per GPU
(src/amuse_ph4/src/gpu.cc:40,56-59 + idata.cc:284-313) happens inside the library here."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from amuse_b200 import g6lib
    return int(g6lib.load().get_device_count())


def test_abi_over_several_devices_matches_oracle_and_one_device():
    ngpu = _ngpu()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    k = 2 if ngpu < 4 else (4 if ngpu < 8 else 8)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multidev_worker.py"), str(k), "20000"],
                         capture_output=True, text=True, timeout=900)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and out.stdout.count("MULTI-DEVICE OK") >= 1, out.stdout[-3000:] + out.stderr[-3000:]


def test_unmodified_ph4_runs_on_several_devices():
    """The reference's own integrator (oracle/_ref/libph4ref_gpu.so = unmodified ph4 -DGPU objects linked to this
    library) with the j-memory spread over all devices: same block steps and energy as on one device."""
    ngpu = _ngpu()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libph4ref_gpu.so")):
        pytest.skip("oracle/_ref/libph4ref_gpu.so not built")
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); from oracle import oracle as O; "
            "from amuse_b200 import plummer as P; m, x, v = P.new_plummer_model(2048, seed=1, do_scale=True); "
            "r = O.ref_evolve(m, x, v, 0.0, 0.14, 0.125, use_gpu=1, libname='libph4ref_gpu.so'); "
            "print('RESULT ' + json.dumps(r))") % ROOT
    import json
    res = {}
    for k in (1, min(ngpu, 8)):
        env = dict(os.environ, G6_B200_DEVICES=str(k))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
        assert out.returncode == 0 and line, out.stdout[-2000:] + out.stderr[-2000:]
        res[k] = json.loads(line[0][7:])
    a, b = res[1], res[min(ngpu, 8)]
    print("ph4 through the ABI on 1 / %d devices: %s / %s" % (min(ngpu, 8), a, b))
    assert abs(a["block_steps"] - b["block_steps"]) <= 0.02 * a["block_steps"]
    assert abs(a["E0"] - b["E0"]) < 1e-9 * abs(a["E0"])
    assert abs((b["E1"] - b["E0"]) / b["E0"]) < 1e-5
