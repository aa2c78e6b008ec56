"""CPU, world_size 2, gloo: the host-side sharding and record exchange of the multi-GPU global path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align(aps):
    for F in (0, 1, 127, 128, 129, 1000, 163840, 10 ** 6):
        for world in (1, 2, 3, 4, 8):
            b = aps.multigpu.shard_bounds(F, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == F
            for (a0, a1), (b0, b1) in zip(b, b[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(q0 % 128 == 0 or q0 == F for q0, _ in b)    # trailing ranks may be empty
            S = aps.multigpu.shard_rows(F, world)
            assert S % 128 == 0 and all(q0 == min(F, r * S) for r, (q0, _) in enumerate(b))   # equal stride
            assert world * S - F < 128 * world                                                 # padding the records need


def _worker(rank, world, port, F, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    mg = ge.load_package().multigpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    bounds = mg.shard_bounds(F, world)
    q0, q1 = bounds[rank]
    rec = torch.full((F + 1024, 2), -7, dtype=torch.int32)   # aps_gplan_records_device layout: (target, partner) rows
    rec[q0:q1, 0] = torch.arange(q0, q1, dtype=torch.int32) % 5            # this rank's targets
    rec[q0:q1, 1] = torch.arange(q0, q1, dtype=torch.int32) + 1            # this rank's partners
    mg.exchange_records(rec, F, world, rank, dist)
    np.save(os.path.join(out_dir, f"rec{rank}.npy"), rec[:F].numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [1000, 128 * 7 + 5])
def test_exchange_records_gloo_world2(tmp_path, F):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() + F) % 2000
    mp.spawn(_worker, args=(2, port, F, str(tmp_path)), nprocs=2, join=True)
    exp = np.stack([np.arange(F) % 5, np.arange(F) + 1], axis=1).astype(np.int32)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"rec{r}.npy"), exp)


def _gather_worker(rank, world, port, counts, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    mg = ge.load_package().multigpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    mats = [rng.standard_normal((c, 8)).astype(np.float32) for c in counts]  # same data on every rank
    pool = torch.full((sum(counts), 8), float("nan"))
    mg.gather_descriptors(pool, mats, rank, world, dist, torch)
    np.save(os.path.join(out_dir, f"pool{rank}.npy"), pool.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[256, 256, 256, 256], [100, 0, 333, 77, 1]])
def test_gather_descriptors_gloo_world2(tmp_path, counts):
    """Each rank uploads its row block, blocks are all-gathered (equal blocks) or broadcast per owner (ragged)."""
    import torch.multiprocessing as mp

    port = 31500 + (os.getpid() + sum(counts)) % 2000
    mp.spawn(_gather_worker, args=(2, port, counts, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    exp = np.vstack([rng.standard_normal((c, 8)).astype(np.float32) for c in counts])
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"pool{r}.npy"), exp)


def _pairlist_worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    pkg = ge.load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full_pp, full_rows, full_met = _fake_pairwise_csr(n)
    # this rank's share: every world-th pair of the column-major pair list, other cells empty
    cnt = np.diff(full_pp)
    keep = np.zeros(n * n, bool)
    o = 0
    for j in range(n):
        for i in range(j):
            keep[i + j * n] = (o % world) == rank
            o += 1
    sel = np.concatenate([np.arange(full_pp[c], full_pp[c + 1]) for c in range(n * n) if keep[c]] + [np.zeros(0, np.int64)]).astype(np.int64)
    pp = np.concatenate([[0], np.cumsum(np.where(keep, cnt, 0))]).astype(np.int64)
    counts, rows_all, met_all = pkg.multigpu.exchange_pairwise_lists(pp, full_rows[sel], full_met[sel], world, dist, torch, "cpu")
    merged = pkg.merge_pairwise_csr(n, counts, rows_all, met_all)
    np.save(os.path.join(out_dir, f"m{rank}.npy"), np.vstack([merged[i][j] for j in range(n) for i in range(j)]))
    np.save(os.path.join(out_dir, f"c{rank}.npy"), np.array([merged[i][j].shape[0] for j in range(n) for i in range(j)]))
    dist.destroy_process_group()


def _fake_pairwise_csr(n):
    rng = np.random.default_rng(n)
    cnt = np.zeros(n * n, np.int64)
    for j in range(n):
        for i in range(j):
            cnt[i + j * n] = rng.integers(0, 6)
    pp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    rows = rng.integers(1, 1000, (int(pp[-1]), 2)).astype(np.uint32)
    return pp, rows, rng.random(int(pp[-1]))


@pytest.mark.parametrize("n", [5, 8])
def test_pairwise_list_exchange_gloo_world2(tmp_path, n):
    """Counts-then-lists all-gather of the per-rank compacted lists reproduces the single-rank cell on every rank."""
    import torch.multiprocessing as mp

    port = 33500 + (os.getpid() + n) % 2000
    mp.spawn(_pairlist_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    pp, rows, _ = _fake_pairwise_csr(n)
    exp_c = np.array([pp[i + j * n + 1] - pp[i + j * n] for j in range(n) for i in range(j)])
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"c{r}.npy"), exp_c)
        assert np.array_equal(np.load(tmp_path / f"m{r}.npy"), rows.astype(np.float64))
