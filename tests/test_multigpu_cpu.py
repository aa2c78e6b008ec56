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
            assert all(q0 % 128 == 0 for q0, _ in b)
            sizes = [q1 - q0 for q0, q1 in b]
            assert max(sizes) - min(sizes) <= 128 or F < 128 * world


def _worker(rank, world, port, F, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    mg = ge.load_package().multigpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    bounds = mg.shard_bounds(F, world)
    q0, q1 = bounds[rank]
    rec = torch.full((2 * F,), -7, dtype=torch.int32)
    rec[q0:q1] = torch.arange(q0, q1, dtype=torch.int32) % 5            # this rank's targets
    rec[F + q0:F + q1] = torch.arange(q0, q1, dtype=torch.int32) + 1    # this rank's partners
    mg.exchange_records(rec, F, bounds, dist)
    np.save(os.path.join(out_dir, f"rec{rank}.npy"), rec.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [1000, 128 * 7 + 5])
def test_exchange_records_gloo_world2(tmp_path, F):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() + F) % 2000
    mp.spawn(_worker, args=(2, port, F, str(tmp_path)), nprocs=2, join=True)
    exp = np.concatenate([np.arange(F) % 5, np.arange(F) + 1]).astype(np.int32)
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
