"""Multi-GPU host logic of the global path: one process per GPU, torch.distributed for the plumbing.

The reference's only parallelism is MATLAB `parfor` over independent work items with the descriptor cell
broadcast to the workers (PP/featureMatching/featureMatchingPairwise.m:54-59, PP/main.m:39-47).  Here:
  * every rank holds ALL descriptors (host upload on rank 0 + NCCL broadcast over NVLink, or its own
    upload), so a query row's k nearest neighbours are complete on the rank that owns the row and no
    cross-GPU top-k merge exists;
  * query rows are sharded in contiguous blocks of 128-row tiles (`shard_bounds`);
  * the only exchange on the data path is the per-query match records (8 bytes per query row:
    int32 target image + uint32 partner index): one in-place all-gather of equal slices (`exchange_records`);
  * compaction into the cell order is deterministic, so the lists are identical for every world size.
Pure index arithmetic + collectives: testable on CPU with the gloo backend (tests/test_multigpu_cpu.py).
"""
from __future__ import annotations

TILE_ROWS = 128


def shard_rows(F: int, world: int) -> int:
    """Rows per rank: the SAME for every rank (a multiple of 128-row tiles), so that one in-place all-gather moves
    every rank's slice; the last ranks' slices may be shorter or empty once clipped to F."""
    blocks = (F + TILE_ROWS - 1) // TILE_ROWS
    return ((blocks + world - 1) // world) * TILE_ROWS


def shard_bounds(F: int, world: int):
    """[(q0, q1)] per rank: contiguous, disjoint, covering [0, F), aligned to 128-row tiles, equal stride."""
    S = shard_rows(F, world)
    return [(min(F, r * S), min(F, (r + 1) * S)) for r in range(world)]


def exchange_records(rec, F: int, world: int, rank: int, dist, group=None):
    """rec: [F + pad, 2] int32 tensor over aps_gplan_records_device (one (target, partner) pair per query row,
    pad >= world * 128).  ONE in-place all-gather of equal slices: after the call every rank holds every rank's
    records (the only exchange on the data path)."""
    S = shard_rows(F, world)
    if world == 1 or S == 0:
        return rec
    if world * S > rec.shape[0]:
        raise ValueError("record buffer too small for equal-sized rank slices")
    dist.all_gather_into_tensor(rec[: world * S], rec[rank * S:(rank + 1) * S], group=group)
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
        exchange_records(rec, plan.F, world, rank, dist)
    plan.compact()


def pairwise_matching_sharded(host, input, allDescriptors, numImg, rank, world, dist=None, ctx=None, torch=None,
                              device=None):
    """featureMatchingPairwise across `world` ranks: rank r computes every world-th image pair of the
    column-major pair list (block-cyclic, like the reference's parfor over the same list,
    featureMatchingPairwise.m:48-59); the compacted per-rank lists are exchanged counts first, then rows and
    metrics (all-gather of per-pair counts, then of the [M x 2] lists at the offsets the counts imply),
    and every rank returns the merged cell.  `host` is the package's host module; `device` is where the exchange
    buffers live ("cuda" for NCCL, "cpu" for gloo)."""
    if world == 1:
        return host.featureMatchingPairwise(input, allDescriptors, numImg, ctx=ctx, shard=(rank, world))
    if torch is None:
        import torch
    n = int(numImg)
    pp, rows, metric = host.featureMatchingPairwise(input, allDescriptors, numImg, ctx=ctx, shard=(rank, world), csr=True)
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    counts, rows_all, met_all = exchange_pairwise_lists(pp, rows, metric, world, dist, torch, dev)
    return host.merge_pairwise_csr(n, counts, rows_all, met_all)


def exchange_pairwise_lists(pair_ptr, rows, metric, world, dist, torch, device, group=None):
    """Counts-then-lists exchange of one rank's compacted match lists.  pair_ptr [n*n + 1] int64, rows [M x 2] uint32,
    metric [M] float64 (host arrays).  Returns (counts [world x n*n] int64, rows [world x Mmax x 2] uint32,
    metric [world x Mmax] float64) as numpy arrays, identical on every rank."""
    import numpy as np

    cells = pair_ptr.size - 1
    cnt = torch.from_numpy(np.diff(pair_ptr).astype(np.int64)).to(device)
    counts = torch.empty((world, cells), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts.view(-1), cnt, group=group)
    m_max = int(counts.sum(dim=1).max().item())
    M = rows.shape[0]
    r_loc = torch.zeros((max(m_max, 1), 2), dtype=torch.int32, device=device)
    m_loc = torch.zeros((max(m_max, 1),), dtype=torch.float64, device=device)
    if M:
        r_loc[:M] = torch.from_numpy(np.ascontiguousarray(rows).view(np.int32)).to(device)
        m_loc[:M] = torch.from_numpy(np.ascontiguousarray(metric)).to(device)
    r_all = torch.empty((world,) + tuple(r_loc.shape), dtype=torch.int32, device=device)
    m_all = torch.empty((world,) + tuple(m_loc.shape), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(r_all.view(-1), r_loc.view(-1), group=group)
    dist.all_gather_into_tensor(m_all.view(-1), m_loc.view(-1), group=group)
    return counts.cpu().numpy(), r_all.cpu().numpy().view(np.uint32), m_all.cpu().numpy()
