"""CPU test of the N>1 path's host logic: world_size-2 (and 3) gloo process groups run the same
sharding + combination code bench.py runs over NCCL.  The per-rank partial forces come from the
oracle here (there is no GPU); what is under test is define_domain, the key packing convention
and combine_partials (sum / min-key / owner-resolved id)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, n, ni, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from amuse_b200 import plummer as P
        from amuse_b200 import sharding as S
        from oracle import oracle as O

        m, x, v = P.new_plummer_model(n, seed=3)
        ids = np.arange(10, 10 + n, dtype=np.int32)
        j0, j1 = S.define_domain(n, world, rank)
        part = O.force(x[:ni], v[:ni], m, x, v, 1e-4, iid=ids[:ni], jid=ids, j_start=j0, j_end=j1)
        d_sum = torch.from_numpy(np.hstack([part["acc"], part["jerk"], -part["pot"][:, None]]).copy())
        r2 = (part["dnn"] ** 2).astype(np.float32)
        key = (r2.view(np.uint32).astype(np.int64) << 32) | part["nn"].astype(np.int64)
        key[part["nn"] < 0] = S.KEY_NONE
        d_key = torch.from_numpy(key)

        def resolve(k):                                   # what g6x_resolve_nn does on the device
            k = k.numpy()
            addr = (k & 0xFFFFFFFF).astype(np.int64)
            own = (k != S.KEY_NONE) & (addr >= j0) & (addr < j1)
            out = np.where(own, ids[np.clip(addr, 0, n - 1)], 0).astype(np.int32)
            if rank == 0:
                out[k == S.KEY_NONE] = -1
            return torch.from_numpy(out)

        nn = S.combine_partials(d_sum, d_key, resolve)
        full = O.force(x[:ni], v[:ni], m, x, v, 1e-4, iid=ids[:ni], jid=ids)
        s = d_sum.numpy()
        ok = (np.allclose(s[:, 0:3], full["acc"], rtol=1e-12, atol=0) and np.allclose(s[:, 3:6], full["jerk"], rtol=1e-10, atol=1e-13)
              and np.allclose(-s[:, 6], full["pot"], rtol=1e-12) and np.array_equal(nn.numpy(), ids[full["nn"]]))
        # neighbour-sphere lists: per-shard lists merged over the ranks == the unsharded list
        h2 = np.minimum(8 * full["dnn"] ** 2, 1.0)
        cnt, lst = [], []
        for i in range(ni):
            c, l = O.neighbours(int(ids[i]), x[i], h2[i], ids[j0:j1], m[j0:j1], x[j0:j1])
            cnt.append(c); lst.append(l)
        tot, merged = S.gather_neighbour_lists(cnt, lst)
        for i in range(ni):
            c, l = O.neighbours(int(ids[i]), x[i], h2[i], ids, m, x)
            ok = ok and tot[i] == c and merged[i] == l.tolist()
        q.put((rank, bool(ok), (j0, j1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_combine_matches_unsharded(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 700, 96, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    doms = sorted(d for _, _, d in res)
    assert doms[0][0] == 0 and doms[-1][1] == 700 and all(doms[k][1] == doms[k + 1][0] for k in range(world - 1))


def test_define_domain_matches_oracle_restatement():
    from amuse_b200 import sharding as S
    from oracle import oracle as O
    for nj in (1, 7, 1024, 1000003):
        for size in (1, 2, 3, 8):
            for rank in range(size):
                a, b = O.define_domain(nj, size, rank)
                assert S.define_domain(nj, size, rank) == (a, min(b, nj))


def test_define_window_tiles_the_address_range_on_tile_boundaries():
    from amuse_b200 import sharding as S
    for nj in (1, 255, 256, 257, 4096, 20000, 1153434):
        for size in (1, 2, 3, 8):
            w = [S.define_window(nj, size, r) for r in range(size)]
            covered = 0
            for lo, hi in w:
                assert lo % 256 == 0 and lo <= hi
                if hi > lo:                      # non-empty windows follow each other without gaps
                    assert lo == covered and hi <= nj
                    covered = hi
            assert covered == nj
