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


# ---- round 2: every row of the bench workload, and BASELINE sizes C3 / C4 / C5 ---------------------------------------
def _knn_tables(aps, ctx, desc, engine, q_ranges=None):
    ctx.set_float_engine(engine)
    try:
        plan = aps.GlobalPlan(ctx, [d.shape[0] for d in desc], desc[0].shape[1], False, 4)
        plan.upload(desc)
        plan.prepare()
        for q0, q1 in (q_ranges or [(0, plan.F)]):
            plan.knn(q0, q1)
        out = [plan.download_knn(q0, q1) for q0, q1 in (q_ranges or [(0, plan.F)])]
        stats = ctx.last_stats()
        plan.close()
    finally:
        ctx.set_float_engine(0)
    return out, stats


@pytest.mark.parametrize("cid", [2, 6])
def test_tensor_engine_equals_exact_engine_on_every_row(aps, cid):
    """The completeness proof of the bf16 tcgen05 path, checked on ALL 163840 rows of the bench workload (and of its
    real-valued twin, config 6): the kNN table must equal, bit for bit, the table of the exact CUDA-core engine
    (set_float_engine(1)), which the small-size tests pin to the oracle."""
    ctx = aps._lib.default_context()
    desc, _ = aps.synth.make_config(cid)
    (tc,), st_tc = _knn_tables(aps, ctx, desc, 2)
    (ex,), st_ex = _knn_tables(aps, ctx, desc, 1)
    assert st_tc["engine"] == "tcgen05" and st_ex["engine"] == "exact"
    assert st_tc["bf16_exact_operands"] == (cid == 2)
    bad = np.flatnonzero((tc[0] != ex[0]).any(1) | (tc[1].view(np.uint32) != ex[1].view(np.uint32)).any(1))
    assert bad.size == 0, f"{bad.size} of {tc[0].shape[0]} rows differ, first {bad[:5]}"


def test_c3_full_size_float_oracle_sample(aps, orc):
    """configs[2]: 100 images x 10000 SIFT-128 (F = 10^6), k=4: 4096 random query rows against the oracle's exact
    search over all 10^6 train rows; every row: table properties; a block of rows: tcgen05 == exact engine."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(3)
    idx, dist, pair_ptr, rows, stats = _run_plan(aps, ctx, desc, False, 4, c["ratio"])
    F = idx.shape[0]
    assert F == 10 ** 6 and stats["engine"] == "tcgen05"
    _check_table_properties(idx, dist)
    X = orc.normalize_rows_global(np.concatenate(desc))
    rng = np.random.default_rng(3)
    sample = np.sort(rng.choice(F, 4096, replace=False))
    oi, od = orc.knn_l2(X, X[sample], 4)
    assert np.array_equal(idx[sample], oi)
    assert np.array_equal(dist[sample].view(np.uint32), od.view(np.uint32))
    # K5 on the full table against the oracle's filter / scatter of the same table
    counts = np.array([d.shape[0] for d in desc], np.int64)
    tgt, par, amb = orc.global_filter(idx, dist, counts, c["ratio"])
    from oracle import oracle as O
    opp = np.zeros(len(desc) ** 2 + 1, np.int64)
    orows = np.zeros((F, 2), np.uint32)
    M = O.lib().orc_global_scatter(tgt, par, F, counts, len(desc), opp, orows.reshape(-1))
    assert amb == 0 and np.array_equal(pair_ptr, opp) and np.array_equal(rows, orows[:M]) and M > 500000
    # 8-way query sharding (what 8 ranks compute) returns the same bits
    idx8, dist8, pp8, rows8, _ = _run_plan(aps, ctx, desc, False, 4, c["ratio"], shards=8)
    assert np.array_equal(idx, idx8) and np.array_equal(dist.view(np.uint32), dist8.view(np.uint32))
    assert np.array_equal(pair_ptr, pp8) and np.array_equal(rows, rows8)
    # exact CUDA-core engine on two blocks of 8192 query rows
    blocks = [(0, 8192), (F - 8192 - 57, F - 57)]
    ex, _ = _knn_tables(aps, ctx, desc, 1, blocks)
    for (q0, q1), (ei, ed) in zip(blocks, ex):
        assert np.array_equal(idx[q0:q1], ei) and np.array_equal(dist[q0:q1].view(np.uint32), ed.view(np.uint32))


def test_c4_full_size_binary(aps, orc):
    """configs[3] at size: 50 images x 20000 ORB 256-bit (F = 10^6), BF Hamming k=4; 4096 sampled rows vs the oracle."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(4)
    idx, dist, pair_ptr, rows, stats = _run_plan(aps, ctx, desc, True, 4, c["ratio"])
    F = idx.shape[0]
    assert F == 10 ** 6
    _check_table_properties(idx, dist)
    X = np.concatenate(desc)
    rng = np.random.default_rng(44)
    sample = np.sort(rng.choice(F, 4096, replace=False))
    oi, od = orc.knn_hamming(X, X[sample], 4)
    assert np.array_equal(idx[sample], oi) and np.array_equal(dist[sample], od)
    counts = np.array([d.shape[0] for d in desc], np.int64)
    tgt, par, _ = orc.global_filter(idx, dist, counts, c["ratio"])
    from oracle import oracle as O
    opp = np.zeros(len(desc) ** 2 + 1, np.int64)
    orows = np.zeros((F, 2), np.uint32)
    M = O.lib().orc_global_scatter(tgt, par, F, counts, len(desc), opp, orows.reshape(-1))
    assert np.array_equal(pair_ptr, opp) and np.array_equal(rows, orows[:M]) and M > 200000


def test_c5_full_size_pairwise_oracle_sample(aps, orc):
    """configs[4] at size: 300 images x 4096 KAZE-64, all 44850 image pairs on the GPU; 240 random image pairs
    (ring neighbours, which hold the planted overlaps, and far pairs) against the oracle's matchFeaturesScratch;
    the 3-shard run (three ranks' shares merged) returns the same cell."""
    ctx = aps._lib.default_context()
    desc, _ = aps.synth.make_config(5)
    n = len(desc)
    inp = {"Matchingmethod": "Exhaustive", "Matchingthreshold": 1.5, "Ratiothreshold": 0.7, "useMATLABFeatureMatch": 0}
    got, met = aps.featureMatchingPairwise(inp, desc, n, ctx=ctx, return_metric=True)
    rng = np.random.default_rng(55)
    pairs = [(i, i + 1) for i in rng.choice(n - 1, 80, replace=False)] + [(i, i + 2) for i in rng.choice(n - 2, 60, replace=False)]
    while len(pairs) < 240:
        i, j = sorted(int(v) for v in rng.choice(n, 2, replace=False))
        pairs.append((i, j))
    total = 0
    for i, j in pairs:
        m, d = orc.match_features(desc[i], desc[j], 1.5, 0.7, True)
        g = got[i][j]
        assert g.shape == (len(m), 2) and np.array_equal(g, m.astype(np.float64)), (i, j)
        if len(m):
            assert np.array_equal(np.asarray(met[i][j]), d), (i, j)
        total += len(m)
    assert total > 20000
    plan = aps.PairwisePlan(ctx, [d.shape[0] for d in desc], 64, False)
    plan.upload(desc)
    plan.prepare()
    shares = [plan.match(1.5, 0.7, r, 3) for r in range(3)]
    plan.close()
    counts = np.stack([np.diff(s[0]) for s in shares])
    m_max = max(s[1].shape[0] for s in shares)
    rows_all = np.zeros((3, m_max, 2), np.uint32)
    for r, s in enumerate(shares):
        rows_all[r, :s[1].shape[0]] = s[1]
    merged = aps.merge_pairwise_csr(n, counts, rows_all)
    for j in range(n):
        for i in range(j):
            assert merged[i][j].shape == got[i][j].shape and np.array_equal(merged[i][j], got[i][j]), (i, j)
