"""Device-resident Hermite integrator (g6x_hermite_evolve): wall seconds per N-body time unit on a Plummer sphere,
optionally with 10 % primordial binaries, on one GPU or -- under torchrun -- with the forces sharded over the ranks
(replicated state, peer-memory exchange; g6x_hermite_set_shard).
Usage: python tools/hermite_timing.py N t_end [eps2] [max_block_steps] [binaries]
       python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/hermite_timing.py 1048576 1.0 0 300 binaries"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import g6lib, plummer as P, sharding as S  # noqa: E402

n = int(sys.argv[1]); t_end = float(sys.argv[2]); eps2 = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
max_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
binaries = len(sys.argv) > 5 and sys.argv[5] == "binaries"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m, x, v = P.new_plummer_model(n, seed=1)
ids = np.arange(1, n + 1, dtype=np.int32)
if binaries:
    ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1)
nt = len(m)
g = g6lib.G6(local)
g.set_j_particles(ids, m, x, v)
if world > 1:
    S.attach_peers(g.L, nt)
    lo, hi = S.define_window(nt, world, rank)
    g.L.g6x_hermite_set_shard(lo, hi)
import time
t0 = time.perf_counter()
g.L.g6x_hermite_init(nt, 0.0, 0.14, eps2, None)
t_init = time.perf_counter() - t0
st = np.zeros(4)
g.L.g6x_hermite_evolve(nt, t_end, 0.14, eps2, max_steps, st)
if rank == 0:
    print("device-resident Hermite N=%d%s eps2=%g on %d GPU(s): init sweep %.3f s (%.3g interactions/s); %.3f s for %.4g time "
          "units = %.4g s per N-body unit; block steps %d, particle steps %d, mean i-block %.1f, %.1f us per block step" % (
              nt, " (10% binaries)" if binaries else "", eps2, world, t_init, float(nt) * nt / t_init, st[3], st[0],
              st[3] / st[0] if st[0] > 0 else float("inf"), st[1], st[2], st[2] / max(1, st[1]), 1e6 * st[3] / max(1, st[1])))
err = g.L.g6x_peer_error() if world > 1 else 0
g.close()
if world > 1:
    dist.destroy_process_group()
sys.exit(1 if err else 0)
