"""Multi-GPU exchange on real devices (needs >= 2 GPUs; skipped on a single-GPU box): the peer-memory
exchange fused into the force kernels (g6x_calc_device_allreduce) against the NCCL collectives and the
FP64 oracle.  The host-side decomposition logic is covered on CPU by tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_memory_exchange_matches_nccl_and_oracle():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else (4 if ngpu < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           "20000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    tail = "\n".join(l for l in (out.stdout + out.stderr).splitlines() if "MULTI-GPU" in l or "g6_b200" in l or
                     "Error" in l or "assert" in l.lower())[-3000:]
    assert out.returncode == 0 and "MULTI-GPU OK" in out.stdout, tail
