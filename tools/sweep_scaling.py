"""BASELINE configs[4]: full i-block force-sweep scaling, N = 64k .. 4M, on the GPUs of this box, against the
FP32 roofline.  Runs bench.py once per N (device-timed value; no e2e / CPU legs) and prints a table.
Usage: python tools/sweep_scaling.py [--gpus G] [--sizes 65536,131072,...]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--sizes", default="65536,131072,262144,524288,1048576,2097152,4194304")
a = ap.parse_args()
print("# full i-block Hermite force sweeps, Plummer, eps2=0, %d x B200 (60 flop/interaction; nominal FP32 peak 74.45 TFLOP/s/GPU)" % a.gpus)
print("# %9s %12s %16s %10s %12s %10s" % ("N", "ms/sweep", "interactions/s", "TFLOP/s", "% of peak", "launches"))
for n in [int(s) for s in a.sizes.split(",")]:
    steps, warm = (1, 3) if n >= 2097152 else (3, 3)
    cmd = [sys.executable]
    if a.gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
                "--master-port", "29533"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--steps", str(steps), "--warmup", str(warm),
            "--particles", str(n), "--no-e2e", "--no-cpu-baseline", "--no-extras", "--parity-sample", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print("# N=%d failed: %s" % (n, out.stderr[-300:]))
        continue
    d = json.loads(lines[-1])
    print("  %9d %12.3f %16.4e %10.2f %12.1f %10d" % (n, d["ms_per_step"], d["value"], d["tflops_60"],
                                                     100 * d["frac_fp32_peak_nominal"], d["gpu_launches"]))
    sys.stdout.flush()
