"""j-domain decomposition over ranks and the cross-rank reduction of partial forces.

Host-side mirror of the only data-parallel strategy the reference has on this path
(SURVEY.md section 2.6): every rank holds the contiguous j-slice of
``jdata::define_domain`` (``src/amuse_ph4/src/jdata.cc:56-67``), computes partial forces on the
whole i-block against it, and the partials are combined like the tail of
``idata::get_acc_and_jerk`` (``src/amuse_ph4/src/idata.cc:284-313``):

    sum    of pot/acc/jerk               -> all_reduce(SUM) on the [ni, 7] double sums
    min    of nearest-neighbour distance -> all_reduce(MIN) on 64-bit keys (float bits of r^2 << 32 | address)
    argmin "only the owner is non-zero"  -> all_reduce(SUM) on the resolved ids

The functions take torch tensors on any device and use whatever process group is active
(NCCL over NVLink on the B200 box, gloo in the CPU tests).  No arithmetic of the force path
lives here.

On the GPU box the same combination is also available INSIDE the library (``attach_peers`` +
``g6x_calc_device_allreduce``): the force kernels store their partials straight into the peers'
exchange buffers over NVLink (CUDA IPC peer memory) while other i-blocks are still being computed,
and one combine kernel per call replaces the three collectives.  ``combine_partials`` (NCCL) stays
as the reference implementation the fused path is tested against.
"""
KEY_NONE = 0x7F800000FFFFFFFF


def define_domain(nj, size, rank):
    """[j_start, j_end) of `rank` -- jdata::define_domain, jdata.cc:56-67."""
    n = nj // size
    if n * size < nj:
        n += 1
    j_start = rank * n
    j_end = j_start + n
    if rank == size - 1:
        j_end = nj
    if j_start >= nj:
        j_end = j_start
    return j_start, min(j_end, nj)


def define_window(nj, size, rank, tile=256):
    """[j_lo, j_hi) of `rank` for the sharded device-resident Hermite step (``g6x_hermite_set_shard``): like
    define_domain, but the windows start on j-tile boundaries (the force kernel's per-tile id ranges and
    FP32 summation groups then coincide with the single-device run, which makes the trajectories
    bit-identical).  Ranks beyond the last tile get an empty window."""
    per = ((nj + size - 1) // size + tile - 1) // tile * tile
    lo = rank * per                       # always a tile boundary, possibly beyond nj (empty window)
    hi = min(nj, (rank + 1) * per)
    return lo, max(lo, hi)


def combine_partials(d_sum, d_key, resolve_ids, group=None):
    """In-place cross-rank combination.

    d_sum : [ni, 7] float64 partial (acc xyz, jerk xyz, sum m/r) of this rank's j-shard
    d_key : [ni]    int64 nearest-neighbour keys of this rank's shard
    resolve_ids(d_key_reduced) -> [ni] int32 tensor: id of the winner if this rank owns its
        address, 0 otherwise, -1 on rank 0 where there is no neighbour at all
        (``g6x_resolve_nn``).
    Returns the [ni] int32 nearest-neighbour ids, identical on every rank.
    """
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return resolve_ids(d_key)
    dist.all_reduce(d_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(d_key, op=dist.ReduceOp.MIN, group=group)
    d_nn = resolve_ids(d_key)
    dist.all_reduce(d_nn, op=dist.ReduceOp.SUM, group=group)
    return d_nn


def attach_peers(lib, capacity, group=None):
    """Set up the library's peer-memory exchange for i-sets of up to `capacity` particles:
    g6x_peer_alloc on every rank, all-gather of the CUDA IPC handles over the process group,
    g6x_peer_attach.  `lib` is the ctypes handle of the opened library (G6.L).  Returns world size."""
    import ctypes as C

    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nbytes = lib.g6x_peer_handle_bytes()
    mine = C.create_string_buffer(nbytes)
    if lib.g6x_peer_alloc(world, rank, int(capacity), mine) != 0:
        raise RuntimeError("g6x_peer_alloc failed (world %d, capacity %d)" % (world, capacity))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(mine.raw), group=group)
    blob = C.create_string_buffer(b"".join(handles), nbytes * world)
    if lib.g6x_peer_attach(blob) != 0:
        raise RuntimeError("g6x_peer_attach failed")
    dist.barrier(group=group)
    return world


def gather_neighbour_lists(counts, lists, group=None):
    """Neighbour-sphere lists of an i-block over all j-shards: every rank passes
    what ``g6_get_neighbour_list_`` returned for ITS shard -- `counts[i]` (full count, may exceed the
    stored length) and `lists[i]` (sorted ids) -- and gets back, identically on every rank,
    (total_counts[i], merged sorted ids[i]).  Ids are unique across shards, so the merge of sorted
    per-shard lists is the list a single device would return (sapporo.cpp:248-272: ascending ids)."""
    import heapq

    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(counts), [list(l) for l in lists]
    world = dist.get_world_size(group)
    mine = ([int(c) for c in counts], [[int(v) for v in l] for l in lists])
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    ni = len(counts)
    tot = [sum(parts[r][0][i] for r in range(world)) for i in range(ni)]
    merged = [list(heapq.merge(*[parts[r][1][i] for r in range(world)])) for i in range(ni)]
    return tot, merged
