"""The C oracle's MATLAB-derived functions against an independent numpy restatement of the same .m lines
(tests/matlab_restatement.py) on random inputs with many ties, and against the golden vectors minted from that
restatement alone (tests/golden/matlab_semantics_v1.npz).  MATLAB is absent, so this is the strongest pin
available for featureMatchingGlobal.m:123-161, matchFeaturesScratch.m:169-215,322-366 and imageMatching.m:75-100."""
import os

import numpy as np
import pytest

import matlab_restatement as mr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cells_equal(oracle_cells, restated_cells):
    """oracle: {(i,j) 0-based: uint32 rows}; restatement: {(i,j) 1-based: float64 rows}"""
    a = {(i + 1, j + 1): v.astype(np.float64) for (i, j), v in oracle_cells.items()}
    assert set(a) == set(restated_cells)
    for key in a:
        assert np.array_equal(a[key], restated_cells[key]), key


def test_normalisations_bitwise(orc):
    rng = np.random.default_rng(1)
    for X in (rng.integers(0, 256, (300, 128)).astype(np.float32), rng.standard_normal((300, 64)).astype(np.float32),
              np.zeros((3, 16), np.float32)):
        assert np.array_equal(orc.normalize_rows_global(X).view(np.uint32), mr.normalize_global(X).view(np.uint32))
        assert np.array_equal(orc.normalize_rows_pairwise(X).view(np.uint32), mr.normalize_pairwise(X).view(np.uint32))


@pytest.mark.parametrize("seed,kind", [(0, "f32"), (1, "f32"), (2, "u8"), (3, "u8")])
def test_global_filter_loop_and_scatter(orc, seed, kind):
    """kNN tables from the oracle's exact search (pinned against OpenCV), filter loop + scatter restated."""
    rng = np.random.default_rng(seed)
    counts = [40, 0, 55, 31, 48]
    if kind == "f32":
        base = rng.integers(0, 6, (80, 16)).astype(np.float32)      # few distinct rows: duplicates and ties everywhere
        desc = [base[rng.integers(0, 80, c)] + (rng.random((c, 16)) < 0.1).astype(np.float32) for c in counts]
    else:
        base = rng.integers(0, 256, (60, 8), dtype=np.uint8)
        desc = [base[rng.integers(0, 60, c)] ^ (rng.random((c, 8)) < 0.15).astype(np.uint8) for c in counts]
    for k, ratio in ((4, 0.8), (4, 1.0), (2, 0.9), (7, 0.6)):
        o = orc.feature_matching_global(desc, k, ratio, return_knn=True)
        cells = mr.global_filter_and_scatter(o["knn_idx"], o["knn_dist"], counts, ratio)
        _cells_equal(o["cells"], cells)  # ratios exactly at the threshold included (compare in single, as MATLAB does)
    assert sum(len(v) for v in cells.values()) > 0


def test_global_k_larger_than_pool(orc):
    rng = np.random.default_rng(5)
    desc = [rng.integers(0, 4, (2, 8)).astype(np.float32), rng.integers(0, 4, (1, 8)).astype(np.float32)]
    o = orc.feature_matching_global(desc, 5, 1.0, return_knn=True)
    assert (o["knn_idx"][:, 3:] == 0).all()                          # missing neighbours: index 0 / +inf
    _cells_equal(o["cells"], mr.global_filter_and_scatter(o["knn_idx"], o["knn_dist"], [2, 1], 1.0))


def test_nearest2_ssd_exact_arithmetic_bitwise(orc):
    """small-integer descriptors: every product and sum is exact in single, so the blocked GEMM form of the
    reference and the oracle's evaluation order must agree bit for bit -- including first-index ties and the
    second minimum with multiplicity."""
    rng = np.random.default_rng(7)
    for (n1, n2, d) in ((200, 150, 32), (64, 1, 8), (50, 333, 16)):
        A = rng.integers(-3, 4, (n1, d)).astype(np.float32)
        B = rng.integers(-3, 4, (n2, d)).astype(np.float32)
        B[: min(n1, n2) // 2] = A[: min(n1, n2) // 2]               # exact duplicates: d1 == 0, ties in d2
        i2, d1, d2 = orc.nearest2_ssd(A, B)
        r2, e1, e2 = mr.nearest2_ssd(A, B)
        assert np.array_equal(i2, r2) and np.array_equal(d1.astype(np.float64), e1) and np.array_equal(d2.astype(np.float64), e2)


def test_nearest2_ssd_real_valued_within_tolerance(orc):
    rng = np.random.default_rng(8)
    A = mr.normalize_pairwise(rng.standard_normal((400, 64)).astype(np.float32))
    B = mr.normalize_pairwise(rng.standard_normal((500, 64)).astype(np.float32))
    i2, d1, d2 = orc.nearest2_ssd(A, B)
    r2, e1, e2 = mr.nearest2_ssd(A, B)
    clear = (e2 - e1) > 1e-5                                          # sgemm summation order may flip closer calls
    assert clear.mean() > 0.99 and np.array_equal(i2[clear], r2[clear])
    assert np.allclose(d1, e1, rtol=0, atol=2e-6) and np.allclose(d2, e2, rtol=0, atol=2e-6)


@pytest.mark.parametrize("unique", [True, False])
def test_match_features_float_end_to_end(orc, unique):
    rng = np.random.default_rng(11)
    A = rng.integers(-2, 3, (300, 24)).astype(np.float32)            # |x| <= 2: no normalisation, exact SSD, many ties
    B = A[rng.permutation(300)[:220]].copy()
    B[rng.random(B.shape) < 0.02] += 1.0
    B = np.clip(B, -2, 2)
    for thr, ratio in ((3.5, 0.6), (100.0, 1.0), (2.0, 0.8)):
        m, met = orc.match_features(A, B, thr, ratio, unique)
        em, emet = mr.match_features(A, B, thr, ratio, unique)
        assert np.array_equal(m, em) and np.array_equal(met, emet.astype(np.float64)), (thr, ratio)
    assert len(em) > 50


def test_match_features_float_normalised_branch(orc):
    """max|x| > 2 -> both sides L2-normalised (eps outside the sqrt).  Real-valued SSDs near 1e-4 carry the
    cancellation error of the GEMM form (~1e-7), which reorders near-equal metrics: same match set, same metrics
    within tolerance, same order wherever consecutive metrics are clearly apart."""
    rng = np.random.default_rng(12)
    A = (rng.random((250, 32)) * 60).astype(np.float32)
    B = A[rng.permutation(250)[:200]] + rng.normal(0, 0.4, (200, 32)).astype(np.float32)
    m, met = orc.match_features(A, B, 1.5, 0.7, True)
    em, emet = mr.match_features(A, B, 1.5, 0.7, True)
    assert len(em) > 100 and set(map(tuple, m)) == set(map(tuple, em))
    assert np.allclose(np.sort(met), np.sort(emet), rtol=0, atol=2e-6)
    pos = {tuple(r): t for t, r in enumerate(em)}
    for t in range(len(m) - 1):
        if met[t + 1] - met[t] > 4e-6:
            assert pos[tuple(m[t])] < pos[tuple(m[t + 1])]


@pytest.mark.parametrize("unique", [True, False])
def test_match_features_binary_end_to_end(orc, unique):
    rng = np.random.default_rng(13)
    A = rng.integers(0, 256, (260, 4), dtype=np.uint8)               # 32-bit descriptors: Hamming ties everywhere
    B = A[rng.permutation(260)[:200]] ^ (rng.random((200, 4)) < 0.1).astype(np.uint8)
    for thr, ratio in ((10.0, 0.8), (100.0, 1.0), (25.0, 0.6)):
        m, met = orc.match_features(A, B, thr, ratio, unique)
        em, emet = mr.match_features(A, B, thr, ratio, unique)
        assert np.array_equal(m, em) and np.array_equal(met, emet.astype(np.float64)), (thr, ratio)
    one = orc.match_features(A, B[:1], 100.0, 1.0, unique)            # N2 == 1: second = nBits
    eone = mr.match_features(A, B[:1], 100.0, 1.0, unique)
    assert np.array_equal(one[0], eone[0]) and np.array_equal(one[1], eone[1].astype(np.float64))


def test_select_partners_and_pack_bits(orc):
    rng = np.random.default_rng(17)
    for n, m in ((1, 6), (2, 6), (3, 1), (9, 4), (30, 6), (64, 3)):
        C = np.triu(rng.integers(0, 4, (n, n)), 1)                    # tiny counts: ties and zero rows
        cand, lin = orc.select_partners(C, m)
        ec, el = mr.select_partners(C, m)
        assert np.array_equal(cand, ec) and np.array_equal(lin + 1, el), (n, m)
    for db in (1, 7, 8, 9, 100, 256):
        bits = (rng.random((13, db)) < 0.5).astype(np.uint8)
        assert np.array_equal(orc.pack_bits(bits), mr.pack_bits(bits))


def test_golden_semantics_vectors(orc):
    """vectors minted from the numpy restatement alone (tests/golden/make_golden_semantics.py)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "matlab_semantics_v1.npz"))
    m, met = orc.match_features(g["mf_A"], g["mf_B"], float(g["mf_thr"]), float(g["mf_ratio"]), True)
    assert np.array_equal(m, g["mf_matches"]) and np.array_equal(met, g["mf_metric"])
    m, met = orc.match_features(g["mb_A"], g["mb_B"], float(g["mb_thr"]), float(g["mb_ratio"]), True)
    assert np.array_equal(m, g["mb_matches"]) and np.array_equal(met, g["mb_metric"])
    counts = g["gl_counts"]
    desc = np.split(g["gl_desc"], np.cumsum(counts)[:-1])
    o = orc.feature_matching_global(desc, int(g["gl_k"]), float(g["gl_ratio"]))
    assert np.array_equal(o["pair_ptr"], g["gl_pair_ptr"]) and np.array_equal(o["rows"], g["gl_rows"])
    counts = g["gf_counts"]
    desc = np.split(g["gf_desc"], np.cumsum(counts)[:-1])
    o = orc.feature_matching_global(desc, int(g["gf_k"]), float(g["gf_ratio"]))
    assert np.array_equal(o["pair_ptr"], g["gf_pair_ptr"]) and np.array_equal(o["rows"], g["gf_rows"]) and len(o["rows"]) > 30
    cand, lin = orc.select_partners(g["sp_counts"], int(g["sp_m"]))
    assert np.array_equal(cand, g["sp_cand"]) and np.array_equal(lin + 1, g["sp_lin"])


@pytest.mark.parametrize("seed", range(12))
def test_randomised_shapes_differential(orc, seed):
    """random small shapes, dimensions, k, thresholds and image counts (including empty images): oracle == restatement"""
    rng = np.random.default_rng(1000 + seed)
    # pairwise matcher, exact-arithmetic floats and short binary codes
    for _ in range(4):
        n1, n2, d = int(rng.integers(1, 90)), int(rng.integers(1, 90)), int(rng.choice([3, 8, 17, 32]))
        A = rng.integers(-2, 3, (n1, d)).astype(np.float32)
        B = rng.integers(-2, 3, (n2, d)).astype(np.float32)
        k = min(n1, n2) // 2
        B[:k] = A[rng.permutation(n1)[:k]]
        thr, ratio, uniq = float(rng.choice([0.5, 3.5, 50.0])), float(rng.choice([0.3, 0.6, 0.9, 1.0])), bool(rng.integers(0, 2))
        m, met = orc.match_features(A, B, thr, ratio, uniq)
        em, emet = mr.match_features(A, B, thr, ratio, uniq)
        assert np.array_equal(m, em) and np.array_equal(met, emet.astype(np.float64)), (n1, n2, d, thr, ratio, uniq)
        nb = int(rng.choice([1, 2, 4]))
        Ab = rng.integers(0, 256, (n1, nb), dtype=np.uint8)
        Bb = rng.integers(0, 256, (n2, nb), dtype=np.uint8)
        Bb[:k] = Ab[rng.permutation(n1)[:k]]
        thr = float(rng.choice([5.0, 25.0, 100.0]))
        m, met = orc.match_features(Ab, Bb, thr, ratio, uniq)
        em, emet = mr.match_features(Ab, Bb, thr, ratio, uniq)
        assert np.array_equal(m, em) and np.array_equal(met, emet.astype(np.float64)), (n1, n2, nb, thr, ratio, uniq)
    # global path on binary codes (exact integer distances): whole function, CSR form
    nimg = int(rng.integers(1, 6))
    counts = [int(c) for c in rng.integers(0, 40, nimg)]
    base = rng.integers(0, 256, (25, 2), dtype=np.uint8)
    desc = [base[rng.integers(0, 25, c)] ^ (rng.random((c, 2)) < 0.1).astype(np.uint8) for c in counts]
    k, ratio = int(rng.choice([2, 3, 4, 6])), float(rng.choice([0.5, 0.8, 1.0]))
    o = orc.feature_matching_global(desc, k, ratio)
    pp, rows = mr.feature_matching_global(desc, k, ratio)
    assert np.array_equal(o["pair_ptr"], pp) and np.array_equal(o["rows"], rows), (counts, k, ratio)
    # partner selection
    n = int(rng.integers(1, 20))
    Cm = np.triu(rng.integers(0, 3, (n, n)), 1)
    msel = int(rng.integers(1, 8))
    cand, lin = orc.select_partners(Cm, msel)
    ec, el = mr.select_partners(Cm, msel)
    assert np.array_equal(cand, ec) and np.array_equal(lin + 1, el)


def test_euclidean_modes_kdtree_subsetpdist2(orc):
    """'kdtree' / 'subsetpdist2' (matchFeaturesScratch.m:142-155, :370-440): Euclidean 2-NN, squared afterwards.
    Independent array-style restatement: pdist2-like distance matrix accumulated column by column in float32."""
    rng = np.random.default_rng(31)
    for N1, N2, D in ((40, 55, 16), (7, 300, 64), (120, 2, 5)):
        A = rng.integers(-8, 9, (N1, D)).astype(np.float32) / 8          # tie-heavy, every sum exact
        B = rng.integers(-8, 9, (N2, D)).astype(np.float32) / 8
        B[0] = A[min(3, N1 - 1)]
        s = np.zeros((N1, N2), np.float32)
        for c in range(D):
            e = A[:, c][:, None] - B[:, c][None, :]
            s = s + e * e
        r = np.sqrt(s)                                                  # euclidean distances, single
        order = np.argsort(r, axis=1, kind="stable")                    # 'Smallest', 2: ascending, first index on ties
        i1 = order[:, 0]
        d1 = (r[np.arange(N1), i1] ** 2).astype(np.float32)             # dBest = d12eu.^2
        d2 = (r[np.arange(N1), order[:, 1]] ** 2).astype(np.float32)
        oi, od1, od2 = orc.nearest2_euclid(A, B)
        assert np.array_equal(oi, (i1 + 1).astype(np.uint32))
        assert np.array_equal(od1.view(np.uint32), d1.view(np.uint32)) and np.array_equal(od2.view(np.uint32), d2.view(np.uint32))
    # through the filters: same lists as feeding the restated distances to the (already cross-checked) filter stage
    A = rng.standard_normal((60, 32)).astype(np.float32)
    B = np.vstack([A[:20] + rng.normal(0, 0.02, (20, 32)).astype(np.float32), rng.standard_normal((50, 32)).astype(np.float32)])
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    for nn in ("kdtree", "subsetpdist2"):
        m, d = orc.match_features_method(A, B, 1.5, 0.7, nn)
        assert len(m) >= 15 and (np.diff(d) >= 0).all() and len(set(m[:, 1].tolist())) == len(m)
    me, _ = orc.match_features_method(A, B, 1.5, 0.7, "exhaustive")
    m0, _ = orc.match_features(A, B, 1.5, 0.7)
    assert np.array_equal(me, m0)


def test_host_method_resolution(aps):
    """Matchingmethod / ApproxFloatNNMethod / useMATLABFeatureMatch of PP/inputs.m:47-49: built modes resolve silently,
    the closed-source branch warns, an unknown method raises (no GPU needed: pure host logic)."""
    import warnings

    from importlib import import_module

    host = import_module(aps.__name__ + ".host")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert host._resolve_method({"ApproxFloatNNMethod": "subsetpdist2"}, "approximate") == 1
        assert host._resolve_method({"ApproxFloatNNMethod": "kdtree", "useMATLABFeatureMatch": 0}, "approximate") == 2
        assert host._resolve_method({}, "exhaustive") == 0
    with pytest.warns(host.ApsSemanticsWarning):
        host._resolve_method({"useMATLABFeatureMatch": 1, "ApproxFloatNNMethod": "subsetpdist2"}, "approximate")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert host._resolve_method({"ApproxFloatNNMethod": "pca2nn"}, "approximate") == 3
        assert host._resolve_method({}, "approximate") == 3               # parser default is pca2nn (:75)
    with pytest.raises(ValueError):
        host._resolve_method({"ApproxFloatNNMethod": "lsh"}, "approximate")


def test_pca2nn_restatement_against_lapack(orc):
    """'pca2nn' (matchFeaturesScratch.m:442-573): the oracle's fixed-arithmetic PCA (float64 covariance + cyclic Jacobi)
    against an independent numpy / LAPACK restatement (eigh of the centred scatter matrix): the similarities depend only
    on the 48-dimensional subspace, so nearest / second-nearest distances agree to rounding and the indices agree wherever
    the two best similarities are not within that rounding."""
    rng = np.random.default_rng(48)
    for N1, N2, D in ((200, 500, 64), (150, 300, 128)):
        A = rng.standard_normal((N1, D)).astype(np.float32)
        B = np.vstack([A[:60] + rng.normal(0, 0.03, (60, D)).astype(np.float32), rng.standard_normal((N2 - 60, D)).astype(np.float32)])
        A /= np.linalg.norm(A, axis=1, keepdims=True)
        B /= np.linalg.norm(B, axis=1, keepdims=True)
        m, d = orc.match_features_pca(A, B, 1.5, 0.8)
        mu = B.mean(0, dtype=np.float64)
        Bc = B.astype(np.float64) - mu
        w, v = np.linalg.eigh(Bc.T @ Bc)
        coeff = v[:, ::-1][:, :48]
        Ap, Bp = (A - mu) @ coeff, Bc @ coeff
        Ap /= np.linalg.norm(Ap, axis=1, keepdims=True)
        Bp /= np.linalg.norm(Bp, axis=1, keepdims=True)
        G = Ap @ Bp.T
        srt = np.sort(G, axis=1)
        i1 = G.argmax(1)
        d1, d2 = 2 - 2 * srt[:, -1], 2 - 2 * srt[:, -2]
        keep = (d1 <= 0.64 * d2) & (d1 <= 1.5)
        assert len(m) >= 50
        q = m[:, 0] - 1
        assert keep[q].all() or (np.abs(d1[q] - 0.64 * d2[q]) < 1e-4).any()
        assert np.array_equal(i1[q] + 1, m[:, 1])
        assert np.allclose(d, d1[q], rtol=0, atol=2e-5)
