"""GPU: BASELINE.json full-size configurations checked through size-independent properties and
through the oracle on a bounded sample (the oracle cannot finish F^2 pairs at these sizes in seconds):
  * kNN table rows of a random query sample == oracle exact kNN of those rows (indices + float bits),
  * every row: ascending distances, (distance, index) order on ties, self present at distance 0,
  * K5 (filter + compaction) on the FULL kNN table == the oracle's filter/scatter of the same table,
  * determinism: a second run returns identical bits; sharded query ranges == one range."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_plan(aps, ctx, desc, is_binary, k, ratio, shards=1):
    counts = [d.shape[0] for d in desc]
    plan = aps.GlobalPlan(ctx, counts, desc[0].shape[1], is_binary, k)
    plan.upload(desc)
    plan.prepare()
    F = plan.F
    for q0, q1 in aps.multigpu.shard_bounds(F, shards):
        plan.knn(q0, q1)
        plan.filter(ratio, q0, q1)
    plan.compact()
    idx, dist = plan.download_knn()
    _, _, pair_ptr, rows = plan.download()
    stats = ctx.last_stats()
    plan.close()
    return idx, dist, pair_ptr, rows, stats


def _check_table_properties(idx, dist):
    F, k = idx.shape
    assert (np.diff(dist, axis=1) >= 0).all()
    tie = np.diff(dist, axis=1) == 0
    assert (np.diff(idx.astype(np.int64), axis=1)[tie] > 0).all()          # ties -> lower index first
    rows = np.arange(1, F + 1, dtype=np.uint32)[:, None]
    self_pos = (idx == rows)
    assert (self_pos.sum(1) == 1).all()
    assert (dist[self_pos] == 0).all()
    assert idx.min() >= 1 and idx.max() <= F


def test_c2_full_size_float(aps, orc):
    """configs[1]: 20 images x 8192 SIFT-128, k=4, ratio 0.8 (the bench workload)."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(2)
    idx, dist, pair_ptr, rows, stats = _run_plan(aps, ctx, desc, False, 4, c["ratio"])
    assert stats["engine"] == "tcgen05"
    F = idx.shape[0]
    assert F == 163840
    _check_table_properties(idx, dist)
    # oracle on a sample of query rows against the full train set
    X = orc.normalize_rows_global(np.concatenate(desc))
    rng = np.random.default_rng(2)
    sample = np.sort(rng.choice(F, 1536, replace=False))
    oi, od = orc.knn_l2(X, X[sample], 4)
    assert np.array_equal(idx[sample], oi)
    assert np.array_equal(dist[sample].view(np.uint32), od.view(np.uint32))
    # K5 on the full table
    counts = np.array([d.shape[0] for d in desc], np.int64)
    tgt, par, amb = orc.global_filter(idx, dist, counts, c["ratio"])
    from oracle import oracle as O
    opp = np.zeros(len(desc) ** 2 + 1, np.int64)
    orows = np.zeros((F, 2), np.uint32)
    M = O.lib().orc_global_scatter(tgt, par, F, counts, len(desc), opp, orows.reshape(-1))
    assert amb == 0
    assert np.array_equal(pair_ptr, opp) and np.array_equal(rows, orows[:M]) and M > 100000
    # determinism + sharding invariance (3 query shards, as three ranks would compute them)
    idx2, dist2, pp2, rows2, _ = _run_plan(aps, ctx, desc, False, 4, c["ratio"], shards=3)
    assert np.array_equal(idx, idx2) and np.array_equal(dist.view(np.uint32), dist2.view(np.uint32))
    assert np.array_equal(pair_ptr, pp2) and np.array_equal(rows, rows2)


def test_c2_query_ranges_with_merged_and_balanced_tails(aps):
    """Query ranges whose unit count leaves a short tail (the scheduler merges the last full round into a balanced
    tail: two lists per row) must return the bits of the one-range run."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(2)
    plan = aps.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
    plan.upload(desc)
    plan.prepare()
    plan.knn()
    idx0, dist0 = plan.download_knn()
    for cuts in ([0, 40448, plan.F], [0, 256 * 228, 256 * 228 + 256 * 300, plan.F]):
        for q0, q1 in zip(cuts[:-1], cuts[1:]):
            plan.knn(q0, q1)
        idx, dist = plan.download_knn()
        assert np.array_equal(idx, idx0) and np.array_equal(dist.view(np.uint32), dist0.view(np.uint32)), cuts
    plan.close()


def test_c4_large_binary(aps, orc):
    """configs[3] family: ORB 256-bit, BF Hamming k=4 (50 images x 4096 here; 20000/img in BASELINE)."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(4, n=50, kp=4096)
    idx, dist, pair_ptr, rows, stats = _run_plan(aps, ctx, desc, True, 4, c["ratio"])
    F = idx.shape[0]
    _check_table_properties(idx, dist)
    X = np.concatenate(desc)
    rng = np.random.default_rng(4)
    sample = np.sort(rng.choice(F, 2048, replace=False))
    oi, od = orc.knn_hamming(X, X[sample], 4)
    assert np.array_equal(idx[sample], oi) and np.array_equal(dist[sample], od)
    counts = np.array([d.shape[0] for d in desc], np.int64)
    tgt, par, _ = orc.global_filter(idx, dist, counts, c["ratio"])
    from oracle import oracle as O
    opp = np.zeros(len(desc) ** 2 + 1, np.int64)
    orows = np.zeros((F, 2), np.uint32)
    M = O.lib().orc_global_scatter(tgt, par, F, counts, len(desc), opp, orows.reshape(-1))
    assert np.array_equal(pair_ptr, opp) and np.array_equal(rows, orows[:M]) and M > 50000


def test_c5_family_kaze64_global_sample(aps, orc):
    """KAZE-64 real-valued descriptors (generic bf16 bound, fallback rows allowed) at 60 x 4096."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(5, n=60, kp=4096)
    idx, dist, pair_ptr, rows, stats = _run_plan(aps, ctx, desc, False, 4, 0.8)
    assert stats["engine"] == "tcgen05"
    F = idx.shape[0]
    _check_table_properties(idx, dist)
    X = orc.normalize_rows_global(np.concatenate(desc))
    rng = np.random.default_rng(5)
    sample = np.sort(rng.choice(F, 1536, replace=False))
    oi, od = orc.knn_l2(X, X[sample], 4)
    assert np.array_equal(idx[sample], oi)
    assert np.array_equal(dist[sample].view(np.uint32), od.view(np.uint32))
    assert stats["fallback_rows"] < 0.01 * F
