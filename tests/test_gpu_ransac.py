"""GPU parity of the RANSAC consumer (csrc/aps_ransac.cu, through the C ABI) against oracle/aps_oracle_ransac.c.
Tolerances: inlier masks, counts, accepted flags, draws consumed: exact; models: 1e-9 relative (the per-trial
arithmetic is bit-identical by construction, the refit uses block-wide double reductions)."""
import numpy as np
import pytest

import ransac_lapack as rl

pytestmark = pytest.mark.gpu
PARAMS = {"maxDistance": 5.5, "inliersConfidence": 99.9, "maxIter": 500}


def _batch(seed, sizes, fracs, n_draws=1000):
    rng = np.random.default_rng(seed)
    P1, P2, ptr, S = [], [], [0], []
    for n, fr in zip(sizes, fracs):
        a, b, _ = rl.make_pair(rng, n, fr)
        P1.append(a), P2.append(b), ptr.append(ptr[-1] + n), S.append(rl.draw_table(rng, n, n_draws))
    return np.array(ptr, np.int64), np.vstack(P1), np.vstack(P2), np.stack(S)


def _compare(g, o, ptr):
    assert np.array_equal(g["draws_used"], o["draws_used"])
    assert np.array_equal(g["n_inliers"], o["n_inliers"])
    assert np.array_equal(g["accepted"], o["accepted"])
    assert np.array_equal(g["inliers"], o["inliers"])
    for p in range(len(ptr) - 1):
        if o["n_inliers"][p] >= 4:
            assert np.allclose(g["models"][p], o["models"][p], rtol=1e-9, atol=1e-12), p
        else:
            assert np.isnan(g["models"][p]).all() and np.isnan(o["models"][p]).all()
        if o["accepted"][p]:
            assert np.allclose(g["models_inv"][p], o["models_inv"][p], rtol=1e-8, atol=1e-12), p
        else:
            assert np.isnan(g["models_inv"][p]).all()


def test_batch_matches_oracle_given_the_same_samples(aps, orc):
    sizes = [60, 200, 500, 1500, 40, 12, 300, 3, 0, 4, 2500]
    fracs = [0.8, 0.5, 0.35, 0.6, 0.2, 1.0, 0.0, 1.0, 0.0, 1.0, 0.45]
    ptr, p1, p2, smp = _batch(5, sizes, fracs)
    o = orc.image_matching_batch(ptr, p1, p2, 5.5, 99.9, 500, smp)
    g = aps.imageMatchingBatch(ptr, p1, p2, PARAMS, samples=smp)
    _compare(g, o, ptr)
    assert o["accepted"].sum() >= 4


@pytest.mark.parametrize("md,conf,mt", [(1.5, 99.0, 100), (3.0, 99.99, 2000), (5.5, 50.0, 20)])
def test_parameters(aps, orc, md, conf, mt):
    ptr, p1, p2, smp = _batch(9, [150, 90, 700], [0.55, 0.3, 0.7], n_draws=2 * mt)
    o = orc.image_matching_batch(ptr, p1, p2, md, conf, mt, smp)
    g = aps.imageMatchingBatch(ptr, p1, p2, {"maxDistance": md, "inliersConfidence": conf, "maxIter": mt}, samples=smp)
    _compare(g, o, ptr)


def test_degenerate_and_invalid_samples(aps, orc):
    rng = np.random.default_rng(3)
    x = np.linspace(0, 500, 50)
    line = np.c_[x, 2 * x + 3]
    same = np.ones((20, 2))
    p1 = np.vstack([line, same])
    p2 = np.vstack([line + 1.0, same])
    ptr = np.array([0, 50, 70], np.int64)
    smp = np.stack([rl.draw_table(rng, 50, 200), np.pad(rl.draw_table(rng, 20, 50), ((0, 150), (0, 0)))])
    o = orc.image_matching_batch(ptr, p1, p2, 5.5, 99.9, 100, smp)
    g = aps.imageMatchingBatch(ptr, p1, p2, {"maxDistance": 5.5, "inliersConfidence": 99.9, "maxIter": 100}, samples=smp)
    _compare(g, o, ptr)
    assert not g["accepted"].any() and not g["inliers"].any()


def test_sample_indices_outside_the_pair_are_skipped(aps, orc):
    ptr, p1, p2, smp = _batch(31, [50, 120], [0.8, 0.6], n_draws=300)
    smp[0, ::3, 2] = 50          # == n: outside
    smp[1, 1::2, 0] = 4_000_000
    o = orc.image_matching_batch(ptr, p1, p2, 5.5, 99.9, 150, smp)
    g = aps.imageMatchingBatch(ptr, p1, p2, {"maxDistance": 5.5, "inliersConfidence": 99.9, "maxIter": 150}, samples=smp)
    _compare(g, o, ptr)
    assert g["accepted"].all()


def test_device_sample_table(aps, orc):
    ptr = np.array([0, 4, 9, 9, 1000, 1003], np.int64)
    t = aps.ransacSampleTable(ptr, 300, seed=42)
    assert t.shape == (5, 300, 4)
    assert np.array_equal(t, aps.ransacSampleTable(ptr, 300, seed=42))
    assert not np.array_equal(t, aps.ransacSampleTable(ptr, 300, seed=43))
    for p, n in enumerate(np.diff(ptr)):
        if n < 4:
            assert not t[p].any()
            continue
        assert t[p].max() < n
        assert all(len(set(r)) == 4 for r in t[p].tolist())
    # roughly uniform: every index of the 991-point pair is drawn, and of the 4-point pair all 4 every time
    assert len(np.unique(t[3])) > 650
    assert np.array_equal(np.sort(t[0], axis=1), np.tile(np.arange(4, dtype=np.uint32), (300, 1)))
    # a run that draws on the device == the oracle fed with the device's table
    ptr2, p1, p2, _ = _batch(17, [300, 80, 900], [0.5, 0.7, 0.3])
    tab = aps.ransacSampleTable(ptr2, 1000, seed=7)
    g = aps.imageMatchingBatch(ptr2, p1, p2, PARAMS, n_draws=1000, seed=7)
    o = orc.image_matching_batch(ptr2, p1, p2, 5.5, 99.9, 500, tab)
    _compare(g, o, ptr2)


def test_image_matching_end_to_end(aps, orc):
    """imageMatching.m mirror: partner selection + device gather + RANSAC + acceptance, against the oracle pieces."""
    n, kp = 8, 600
    keypoints, matches, truth = aps.synth.synth_matched_keypoints(n, kp, seed=123)
    inp = dict(PARAMS, mBrownLowe=4)
    counts = np.array([[np.asarray(matches[i][j]).shape[0] if np.asarray(matches[i][j]).ndim == 2 else 0
                        for j in range(n)] for i in range(n)])
    _, lin = orc.select_partners(counts, 4)
    ptr = np.concatenate([[0], np.cumsum([counts[c % n, c // n] for c in lin])]).astype(np.int64)
    tab = aps.ransacSampleTable(ptr, 1000, seed=5)
    allM, numM, tf = aps.imageMatching(inp, n, keypoints, matches, seed=5)
    # oracle: gather like refineMatch, then the batch
    P1 = np.vstack([keypoints[c // n][np.asarray(matches[c % n][c // n], np.int64)[:, 1] - 1] for c in lin])
    P2 = np.vstack([keypoints[c % n][np.asarray(matches[c % n][c // n], np.int64)[:, 0] - 1] for c in lin])
    o = orc.image_matching_batch(ptr, P1, P2, 5.5, 99.9, 500, tab)
    n_acc = 0
    for p, c in enumerate(lin):
        i, j = int(c % n), int(c // n)
        if o["accepted"][p]:
            n_acc += 1
            want = np.asarray(matches[i][j])[o["inliers"][ptr[p]:ptr[p + 1]]]
            assert np.array_equal(allM[i][j], want)
            assert numM[i, j] == o["n_inliers"][p]
            assert np.allclose(tf[i][j], o["models"][p], rtol=1e-9, atol=1e-12)
            assert np.allclose(tf[j][i], o["models_inv"][p], rtol=1e-8, atol=1e-12)
            if (i, j) in truth:  # and the model is the planted homography (image j -> image i)
                H = truth[(i, j)]
                assert np.allclose(tf[i][j] / tf[i][j][2, 2], H / H[2, 2], rtol=2e-2, atol=2.0)
        else:
            assert allM[i][j].size == 0 and numM[i, j] == 0 and tf[i][j] is None and tf[j][i] is None
    assert n_acc >= n  # every ring neighbour pair is recovered
    assert np.count_nonzero(numM) == n_acc


def test_c2_like_scale_and_timing(aps, orc, capsys):
    """C2-like size (20 images x 8192 keypoints, m = 6: 76 candidate pairs, ~58 k correspondences): identical inlier
    sets at full size, and both sides timed on the same input (GPU end to end incl. copies vs oracle on all host cores)."""
    import time

    n, kp = 20, 8192
    keypoints, matches, _ = aps.synth.synth_matched_keypoints(n, kp, seed=77)
    inp = dict(PARAMS, mBrownLowe=6)
    aps.imageMatching(inp, n, keypoints, matches, seed=1)  # warm-up (allocations, module load)
    t0 = time.perf_counter()
    aps.imageMatching(inp, n, keypoints, matches, seed=1)
    t_gpu = time.perf_counter() - t0
    last = aps.imageMatching.last
    lin, ptr = last["pairs_lin"], last["pt_ptr"]
    P1 = np.vstack([keypoints[c // n][np.asarray(matches[c % n][c // n], np.int64)[:, 1] - 1] for c in lin])
    P2 = np.vstack([keypoints[c % n][np.asarray(matches[c % n][c // n], np.int64)[:, 0] - 1] for c in lin])
    tab = aps.ransacSampleTable(ptr, 1000, seed=1)
    t0 = time.perf_counter()
    o = orc.image_matching_batch(ptr, P1, P2, 5.5, 99.9, 500, tab)
    t_cpu = time.perf_counter() - t0
    assert np.array_equal(o["inliers"], last["inliers"]) and np.array_equal(o["accepted"], last["accepted"])
    assert np.array_equal(o["draws_used"], last["draws_used"]) and o["accepted"].sum() >= 2 * n
    assert np.allclose(last["models"][o["accepted"]], o["models"][o["accepted"]], rtol=1e-9, atol=1e-12)
    with capsys.disabled():
        print(f"\n[ransac C2-like] {len(lin)} pairs, {int(ptr[-1])} correspondences: GPU {t_gpu * 1e3:.1f} ms end to end, "
              f"oracle {t_cpu * 1e3:.1f} ms on {orc.num_threads()} host threads")


def test_gpu_against_committed_lapack_vectors(aps):
    """The GPU path against tests/golden/ransac_v1.npz (numpy/LAPACK restatement; no oracle involved)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ransac_v1.npz"))
    for c in [str(x) for x in g["cases"]]:
        md, conf, mt = g[f"{c}_params"]
        inp = {"maxDistance": md, "inliersConfidence": conf, "maxIter": int(mt)}
        m, inl, found = aps.estimateTransformationRANSAC(g[f"{c}_p1"], g[f"{c}_p2"], "projective", inp,
                                                         samples=g[f"{c}_samples"])
        assert found, c
        assert np.array_equal(inl, g[f"{c}_inliers"]), c
        gm = g[f"{c}_model"]
        assert np.allclose(m / m[2, 2], gm / gm[2, 2], rtol=1e-7, atol=1e-9), c


def test_reference_error_behaviour(aps):
    kps = [np.zeros((5, 2)), np.zeros((5, 2))]
    bad = [[np.zeros((0, 0)), np.array([[1.0, 9.0], [2, 2], [3, 3], [4, 4]])], [np.zeros((0, 0)), np.zeros((0, 0))]]
    with pytest.raises(aps.ApsError) as e:
        aps.imageMatching(dict(PARAMS, mBrownLowe=1), 2, kps, bad)
    assert e.value.identifier == "refineMatch:MatchIndexOutOfBounds"
    with pytest.raises(ValueError):
        aps.imageMatching(PARAMS, 3, kps, bad)
    with pytest.raises(ValueError):
        aps.estimateTransformationRANSAC(np.zeros((4, 2)), np.zeros((5, 2)))
    m, inl, found = aps.estimateTransformationRANSAC(np.zeros((3, 2)), np.zeros((3, 2)), "projective", PARAMS)
    assert m is None and not found and inl.shape == (3,) and not inl.any()
