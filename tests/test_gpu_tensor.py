"""GPU: the tcgen05 candidate kernel in isolation (raw scores against a float64 matmul of the
bf16-rounded operands) and the tensor engine end to end (forced) against the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16_round(x):
    """float32 -> nearest-even bfloat16, returned as float32."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def check_candidates(scores, cidx, csc, lo, hi):
    """The two lists of a segment cover disjoint tile subsets: their union must contain the segment's
    8 best columns (as the kernel's own epilogue values rank them) and carry the kernel's scores."""
    seg = scores[:, lo:hi]
    k8 = min(8, hi - lo)
    best = np.argsort(-seg, axis=1, kind="stable")[:, :k8] + lo
    for r in range(scores.shape[0]):
        valid = cidx[r] != 0xFFFFFFFF
        cand = cidx[r][valid].astype(np.int64)
        assert len(set(cand.tolist())) == len(cand) and ((cand >= lo) & (cand < hi)).all(), r
        kth = seg[r, best[r, -1] - lo]
        need = set(np.flatnonzero(seg[r] > kth) + lo)      # strictly better than the 8th: must be present
        assert need <= set(cand.tolist()), r
        assert len(cand) >= k8, r
        # retained scores carry their slot number in the 3 low mantissa bits (kernel-internal packing)
        assert np.array_equal(csc[r][valid].view(np.uint32) & ~np.uint32(7),
                              scores[r, cand].view(np.uint32) & ~np.uint32(7)), r


def tc_scores(aps, Q, T, nseg=1, want_scores=True):
    ctx = aps._lib.default_context()
    L = aps._lib.lib()
    Q = np.ascontiguousarray(Q, np.float32)
    T = np.ascontiguousarray(T, np.float32)
    nq, nt = Q.shape[0], T.shape[0]
    scores = np.zeros((nq, nt), np.float32) if want_scores else None
    slots = 8 * L.aps_debug_tc_slots(ctx.handle, nq, nt)   # lists of 8 per row (one per column segment of its unit)
    cidx = np.zeros((nq, 1, slots), np.uint32)
    csc = np.zeros((nq, 1, slots), np.float32)
    aps._lib.check(L.aps_debug_tc_scores(ctx.handle, Q.ctypes.data, nq, T.ctypes.data, nt, Q.shape[1], nseg,
                                         scores.ctypes.data if want_scores else None, cidx.ctypes.data, csc.ctypes.data))
    return scores, cidx, csc


@pytest.mark.parametrize("nq,nt,D", [(128, 256, 128), (100, 300, 128), (257, 1000, 64), (130, 513, 100), (64, 40, 128)])
def test_tc_raw_scores_match_bf16_matmul(aps, nq, nt, D):
    rng = np.random.default_rng(nq * 7 + nt)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    T = rng.standard_normal((nt, D)).astype(np.float32)
    got, cidx, csc = tc_scores(aps, Q, T)
    Qb, Tb = bf16_round(Q).astype(np.float64), bf16_round(T).astype(np.float64)
    sq = np.zeros(nt, np.float32)
    for d in range(D):  # sequential float32 sum of squares, as K1 computes it
        sq = sq + T[:, d] * T[:, d]
    exp = Qb @ Tb.T - 0.5 * sq.astype(np.float64)[None, :]
    err = np.abs(got - exp).max()
    assert err < 2e-3, err          # fp32 accumulation of 128 products of magnitude ~1
    check_candidates(got, cidx[:, 0], csc[:, 0], 0, nt)


def test_tc_integer_descriptors_are_exact(aps):
    """0..255 integer descriptors are exact in bf16 and their dot products exact in fp32 accumulate."""
    rng = np.random.default_rng(3)
    Q = rng.integers(0, 256, (256, 128)).astype(np.float32)
    T = rng.integers(0, 256, (512, 128)).astype(np.float32)
    got, _, _ = tc_scores(aps, Q, T)
    sq = (T.astype(np.float64) ** 2).sum(1)
    exp = Q.astype(np.float64) @ T.astype(np.float64).T - 0.5 * sq[None, :]
    assert np.array_equal(got.astype(np.float64), exp.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("nq,nt", [(300, 5000), (256 * 150, 700), (256 * 149 + 10, 1300), (256 * 228, 1536),
                                   (256 * 208 - 77, 2000), (256 * 158, 1536)])
def test_tc_tail_units_split_columns(aps, nq, nt):
    """Work units of the last partial round are split into column segments (several lists per row);
    whatever the split, the union of a row's lists must contain its best columns."""
    rng = np.random.default_rng(nq % 1000 + nt)
    Q = rng.standard_normal((nq, 64)).astype(np.float32)
    T = rng.standard_normal((nt, 64)).astype(np.float32)
    _, cidx, csc = tc_scores(aps, Q, T, want_scores=False)
    Qb, Tb = bf16_round(Q).astype(np.float64), bf16_round(T).astype(np.float64)
    sq = np.zeros(nt, np.float32)
    for d in range(64):
        sq = sq + T[:, d] * T[:, d]
    for r in list(range(0, nq, max(1, nq // 97))) + [nq - 1]:
        sc = Qb[r] @ Tb.T - 0.5 * sq
        top = np.argsort(-sc, kind="stable")[:6]
        cand = set(cidx[r, 0][cidx[r, 0] != 0xFFFFFFFF].tolist())
        assert set(top[:5].tolist()) <= cand, r       # float32-vs-float64 near-ties aside, the best are there


@pytest.mark.parametrize("cid,n,kp", [(1, 6, 2048), (5, 6, 1500)])
def test_tensor_engine_global_vs_oracle(aps, orc, cid, n, kp):
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(cid, n=n, kp=kp)
    ref = orc.feature_matching_global(desc, 4, c["ratio"], return_knn=True)
    ctx.set_float_engine(2)
    try:
        counts = [d.shape[0] for d in desc]
        plan = aps.GlobalPlan(ctx, counts, desc[0].shape[1], False, 4)
        plan.upload(desc)
        plan.prepare()
        plan.knn()
        plan.filter(c["ratio"])
        plan.compact()
        matches, _, pair_ptr, rows = plan.download()
        stats = ctx.last_stats()
        F = plan.F
        idx, dist = plan.download_knn()
        plan.close()
    finally:
        ctx.set_float_engine(0)
    assert stats["engine"] == "tcgen05"
    assert np.array_equal(idx, ref["knn_idx"])
    assert np.array_equal(dist.view(np.uint32), ref["knn_dist"].view(np.uint32))
    assert np.array_equal(pair_ptr, ref["pair_ptr"]) and np.array_equal(rows, ref["rows"])
    print("tensor engine stats", cid, stats)
    if cid == 1:
        assert stats["bf16_exact_operands"] and stats["fallback_rows"] < 0.02 * F


def test_tensor_engine_pairwise_and_knn_entry(aps, orc):
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(5, n=3, kp=2100)
    ctx.set_float_engine(2)
    try:
        m, met = aps.matchFeaturesScratch(desc[0], desc[1], MatchThreshold=1.5, MaxRatio=0.7)
        X = orc.normalize_rows_global(np.concatenate(desc))
        idx, dist = aps.flann_knn_win(X, X[:1000].copy(), 4)
    finally:
        ctx.set_float_engine(0)
    om, omet = orc.match_features(desc[0], desc[1], 1.5, 0.7)
    assert np.array_equal(m, om) and np.array_equal(met, omet) and len(m) > 100
    oi, od = orc.knn_l2(X, X[:1000], 4)
    assert np.array_equal(idx, oi) and np.array_equal(dist.view(np.uint32), od.view(np.uint32))


def test_tensor_engine_unprovable_rows_fall_back_to_exact_search(aps, orc):
    """Rows whose candidate set cannot be proven complete (here: 40 identical descriptors, so the k-th
    distance ties with the worst retained candidate) are re-searched by the exact kernel."""
    ctx = aps._lib.default_context()
    rng = np.random.default_rng(17)
    X = rng.standard_normal((6000, 128)).astype(np.float32)
    X[100:140] = X[100]                       # 40 exact duplicates
    X[3000:3005] = X[100]
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    ctx.set_float_engine(2)
    try:
        idx, dist = aps.flann_knn_win(X, X, 4)
        stats = ctx.last_stats()
    finally:
        ctx.set_float_engine(0)
    oi, od = orc.knn_l2(X, X, 4)
    assert stats["engine"] == "tcgen05" and stats["fallback_rows"] >= 40
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))


def test_tensor_engine_real_valued_sift_second_pass(aps, orc):
    """Real-valued SIFT-like descriptors (not exact in bf16, neighbour distances tightly packed): many rows
    cannot be proven with 8 candidates and take the second tensor pass (32 candidates), a few the exact
    kernel; the result must still be the oracle's, bit for bit."""
    ctx = aps._lib.default_context()
    desc, c = aps.synth.make_config(6, n=8, kp=2048)
    ref = orc.feature_matching_global(desc, 4, c["ratio"], return_knn=True)
    ctx.set_float_engine(2)
    try:
        plan = aps.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
        plan.upload(desc)
        plan.prepare()
        plan.knn()
        plan.filter(c["ratio"])
        plan.compact()
        _, _, pair_ptr, rows = plan.download()
        idx, dist = plan.download_knn()
        stats = ctx.last_stats()
        plan.close()
    finally:
        ctx.set_float_engine(0)
    assert stats["engine"] == "tcgen05" and not stats["bf16_exact_operands"]
    assert np.array_equal(idx, ref["knn_idx"])
    assert np.array_equal(dist.view(np.uint32), ref["knn_dist"].view(np.uint32))
    assert np.array_equal(pair_ptr, ref["pair_ptr"]) and np.array_equal(rows, ref["rows"])


def test_exact_operand_variant_still_proves_completeness(aps, orc):
    """Integer-valued descriptors are exact in the tensor operands and take the 6-candidate variant of the candidate
    kernel (lists stay 8 wide, two slots unused).  The unused slots must not read as "every column was a candidate":
    rows with more exact duplicates than a list holds cannot be proven from 6 candidates (the k-th distance ties with
    the worst retained one), so they must take the second pass / the exact engine -- and the result is the oracle's."""
    ctx = aps._lib.default_context()
    rng = np.random.default_rng(23)
    desc = [np.rint(np.abs(rng.standard_normal((3000, 128))) * 60).clip(0, 255).astype(np.float32) for _ in range(3)]
    desc[0][200:212] = desc[0][200]           # 12 copies inside one image
    desc[1][5:45] = desc[0][200]              # 40 more in another: more than the 32 candidates of the second pass
    desc[2][7] = desc[0][200]
    ref = orc.feature_matching_global(desc, 4, 0.8, return_knn=True)
    ctx.set_float_engine(2)
    try:
        plan = aps.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
        plan.upload(desc)
        plan.prepare()
        plan.knn()
        idx, dist = plan.download_knn()
        stats = ctx.last_stats()
        first = ctx.first_pass_unproven()
        plan.close()
    finally:
        ctx.set_float_engine(0)
    assert stats["engine"] == "tcgen05" and stats["bf16_exact_operands"]
    assert first >= 53 and stats["fallback_rows"] >= 53          # the 53 duplicate rows at least
    assert np.array_equal(idx, ref["knn_idx"])
    assert np.array_equal(dist.view(np.uint32), ref["knn_dist"].view(np.uint32))
