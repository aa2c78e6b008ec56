"""Multi-GPU host logic of the global path: one process per GPU, torch.distributed for the plumbing.

The reference's only parallelism is MATLAB `parfor` over independent work items with the descriptor cell
broadcast to the workers (PP/featureMatching/featureMatchingPairwise.m:54-59, PP/main.m:39-47).  Here:
  * every rank holds ALL descriptors (host upload on rank 0 + NCCL broadcast over NVLink, or its own
    upload), so a query row's k nearest neighbours are complete on the rank that owns the row and no
    cross-GPU top-k merge exists;
  * query rows are sharded in contiguous blocks of 128-row tiles (`shard_bounds`);
  * the only exchange on the data path is the per-query match records (8 bytes per query row:
    int32 target image + uint32 partner index), each rank broadcasting its slice (`exchange_records`);
  * compaction into the cell order is deterministic, so the lists are identical for every world size.
Pure index arithmetic + collectives: testable on CPU with the gloo backend (tests/test_multigpu_cpu.py).
"""
from __future__ import annotations

TILE_ROWS = 128


def shard_bounds(F: int, world: int):
    """[(q0, q1)] per rank: contiguous, disjoint, covering [0, F), aligned to 128-row tiles."""
    blocks = (F + TILE_ROWS - 1) // TILE_ROWS
    out = []
    for r in range(world):
        b0, b1 = r * blocks // world, (r + 1) * blocks // world
        out.append((min(F, b0 * TILE_ROWS), min(F, b1 * TILE_ROWS)))
    return out


def exchange_records(rec, F: int, bounds, dist, group=None):
    """rec: 1-D int32 tensor of 2*F entries (target[F] then partner[F], the layout of
    aps_gplan_records_device).  After the call every rank holds every rank's slice."""
    for r, (a, b) in enumerate(bounds):
        if b > a:
            dist.broadcast(rec[a:b], src=r, group=group)
            dist.broadcast(rec[F + a:F + b], src=r, group=group)
    return rec


def upload_row_block(pool, host_mats, q0: int, q1: int, torch):
    """Copies rows [q0, q1) of the pooled descriptor matrix from the per-image host matrices (pinned numpy views
    or tensors, image i = rows off[i]..off[i+1]) into `pool` ([F x D] tensor, device or CPU).  Asynchronous on the
    current stream when the sources are pinned."""
    off = 0
    for m in host_mats:
        n = int(m.shape[0])
        a, b = max(q0, off), min(q1, off + n)
        if b > a:
            src = m if torch.is_tensor(m) else torch.from_numpy(m)
            pool[a:b].copy_(src[a - off:b - off], non_blocking=True)
        off += n
    return pool


def gather_descriptors(pool, host_mats, rank: int, world: int, dist, torch, group=None):
    """Descriptor hand-over for one process per GPU: every rank uploads ITS block of rows (shard_bounds) from host
    memory, then the blocks are all-gathered over NVLink, so the host->device traffic is spread over all the
    ranks' PCIe links instead of one (ncclAllGather of descriptor blocks).  pool: [F x D]."""
    F = int(pool.shape[0])
    bounds = shard_bounds(F, world)
    q0, q1 = bounds[rank]
    upload_row_block(pool, host_mats, q0, q1, torch)
    if world == 1:
        return pool
    sizes = {b - a for a, b in bounds}
    if len(sizes) == 1 and q1 > q0:  # equal blocks: one in-place all-gather
        dist.all_gather_into_tensor(pool.reshape(-1), pool[q0:q1].reshape(-1), group=group)
    else:
        for r, (a, b) in enumerate(bounds):
            if b > a:
                dist.broadcast(pool[a:b], src=r, group=group)
    return pool


class CudaView:
    """Exposes a raw device pointer to torch via __cuda_array_interface__ (zero-copy, for collectives)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def global_matching_step(plan, ratio, rank, world, dist=None, rec=None):
    """One sharded pass of the global pipeline on an uploaded/broadcast plan (all ranks call it)."""
    bounds = shard_bounds(plan.F, world)
    q0, q1 = bounds[rank]
    plan.prepare()
    plan.knn(q0, q1)
    plan.filter(ratio, q0, q1)
    if world > 1:
        exchange_records(rec, plan.F, bounds, dist)
    plan.compact()


def pairwise_matching_sharded(host, input, allDescriptors, numImg, rank, world, dist=None, ctx=None):
    """featureMatchingPairwise across `world` ranks: rank r computes every world-th image pair of the
    column-major pair list (block-cyclic, like the reference's parfor over the same list,
    featureMatchingPairwise.m:48-59), the per-pair lists are exchanged and every rank returns the merged cell.
    `host` is the package's host module (featureMatchingPairwise / merge_pairwise_shards)."""
    mine = host.featureMatchingPairwise(input, allDescriptors, numImg, ctx=ctx, shard=(rank, world))
    if world == 1:
        return mine
    shards = [None] * world
    dist.all_gather_object(shards, mine)  # match lists are small next to the descriptors (8-16 B per match)
    return host.merge_pairwise_shards(shards)
