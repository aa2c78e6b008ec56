"""GPU: stage 1 of the batched pairwise path (csrc/aps_pair_screen.cu) -- fp16 tensor-core screen of every
(query row, train image).  Checked here: the accumulator layout the epilogue assumes, the error bound the rejection
proof uses (measured error must stay below HALF of aps_pair_screen_dot_eps), the two values kept per row, and that the
match lists with the screen on are the lists with the screen off (which the other tests pin to the oracle)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dot_eps(D):
    Dp = (D + 63) // 64 * 64
    return (Dp // 16) * 2.0 ** -10 + 1.5e-3          # aps_pair_screen_dot_eps (csrc/aps_common.cuh)


def _screen(aps, A, B, dump_tiles):
    L = aps._lib.lib()
    ctx = aps._lib.default_context()
    A, B = np.ascontiguousarray(A, np.float32), np.ascontiguousarray(B, np.float32)
    out = np.zeros((A.shape[0], 2), np.float32)
    dump = np.zeros((A.shape[0], dump_tiles, 64), np.uint32)
    aps._lib.check(L.aps_debug_pair_screen(ctx.handle, A.ctypes.data, A.shape[0], B.ctypes.data, B.shape[0], A.shape[1],
                                           out.ctypes.data, dump.ctypes.data, dump_tiles))
    return out, dump


def _cases(rng, N1, N2, D):
    def unit(x):
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    yield "unit gaussian", unit(rng.standard_normal((N1, D))), unit(rng.standard_normal((N2, D)))
    yield "positive, |.| <= 2 (monotone partial sums)", rng.uniform(0, 2, (N1, D)), rng.uniform(0, 2, (N2, D))
    yield "mixed magnitudes", rng.standard_normal((N1, D)) * rng.uniform(1e-3, 0.6, (N1, 1)), rng.uniform(-2, 2, (N2, D))
    s = unit(np.abs(rng.standard_normal((N1, D))))
    yield "near duplicates", s, np.vstack([s + rng.normal(0, 0.02, s.shape), unit(np.abs(rng.standard_normal((N2 - N1, D))))])


@pytest.mark.parametrize("D", [64, 128, 61])
def test_accumulator_layout_and_error_bound(aps, D):
    rng = np.random.default_rng(D)
    N1, N2 = 256, 1000                       # train rows start at pooled row 256: tile-aligned, last tile partial
    tiles = (N2 + 127) // 128
    for name, A, B in _cases(rng, N1, N2, D):
        A, B = A.astype(np.float32), B.astype(np.float32)
        out, dump = _screen(aps, A, B, tiles)
        S = dump.view(np.float16).reshape(N1, tiles * 128)[:, :N2].astype(np.float64)    # element 2j / 2j+1 = lo / hi half
        exact = A.astype(np.float64) @ B.astype(np.float64).T
        scale = np.linalg.norm(A.astype(np.float64), axis=1)[:, None] * np.linalg.norm(B.astype(np.float64), axis=1)[None, :]
        err = np.abs(S - exact) / np.maximum(scale, 1e-30)
        print(f"D={D} {name}: max |fp16 tensor dot - exact| / (|a||b|) = {err.max():.2e} (bound {dot_eps(D):.2e})")
        assert err.max() <= 0.5 * dot_eps(D), (name, err.max(), dot_eps(D))
        # the two values kept per row: best over all columns, second of the 8 column classes (column mod 8)
        cls = np.full((N1, 8), -65504.0)
        for c in range(8):
            cls[:, c] = S[:, c::8].max(axis=1)
        srt = np.sort(cls, axis=1)
        assert np.array_equal(out[:, 0].astype(np.float64), srt[:, -1]), name
        assert np.array_equal(out[:, 1].astype(np.float64), srt[:, -2]), name


def test_partial_tiles_are_masked(aps):
    """Train image starting and ending inside a tile: foreign columns (the query rows themselves: dot(a, a) = 1 would
    win) must never be selected."""
    rng = np.random.default_rng(9)
    D, N1, N2 = 64, 300, 777
    A = rng.standard_normal((N1, D)).astype(np.float32)
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    B = rng.standard_normal((N2, D)).astype(np.float32)
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    out, _ = _screen(aps, A, B, 1)
    dots = (A.astype(np.float16).astype(np.float64) @ B.astype(np.float16).astype(np.float64).T)
    best = dots.max(axis=1)
    assert np.abs(out[:, 0] - best).max() <= dot_eps(D)      # a masked-in foreign column would show here
    assert (out[:, 0] < 0.9).all() and (out[:, 1] <= out[:, 0]).all()


def test_screen_on_equals_screen_off(aps, orc):
    ctx = aps._lib.default_context()
    inp = {"Matchingmethod": "Exhaustive", "Matchingthreshold": 1.5, "Ratiothreshold": 0.7, "useMATLABFeatureMatch": 0}
    for cid, n, kp in ((5, 24, 1500), (5, 7, 4096), (1, 5, 2048), (6, 5, 1200)):
        desc, _ = aps.synth.make_config(cid, n=n, kp=kp)
        ctx.set_pairwise_screen(True)
        on, mon = aps.featureMatchingPairwise(inp, desc, len(desc), ctx=ctx, return_metric=True)
        st = ctx.pairwise_stats()
        ctx.set_pairwise_screen(False)
        try:
            off, moff = aps.featureMatchingPairwise(inp, desc, len(desc), ctx=ctx, return_metric=True)
        finally:
            ctx.set_pairwise_screen(True)
        npairs = sum(1 for j in range(len(desc)) for i in range(j) if desc[i].shape[0] and desc[j].shape[0])
        print(f"config {cid} n={n} kp={kp}: {st}")
        assert st["pairs_screened"] == npairs and st["pairs_to_exact"] <= npairs
        if cid == 5 and n == 24:
            assert st["pairs_to_exact"] < 0.4 * npairs       # only ring neighbours hold planted overlaps
        for j in range(len(desc)):
            for i in range(j):
                assert on[i][j].shape == off[i][j].shape and np.array_equal(on[i][j], off[i][j]), (cid, i, j)
                if on[i][j].shape[0]:
                    assert np.array_equal(np.asarray(mon[i][j]), np.asarray(moff[i][j]))
        if n <= 7:   # and against the oracle on every pair of the small sets
            ref = orc.feature_matching_pairwise(desc, 1.5, 0.7)["cells"]
            for j in range(len(desc)):
                for i in range(j):
                    exp = ref.get((i, j))
                    assert (exp is None and on[i][j].shape[0] == 0) or np.array_equal(on[i][j], exp.astype(np.float64)), (cid, i, j)


@pytest.mark.parametrize("nn", ["subsetpdist2", "kdtree"])
def test_approximate_float_methods_match_the_oracle(aps, orc, nn):
    """Matchingmethod='Approximate' (the reference's inputs.m default, with 'subsetpdist2'): the two float modes that
    are Euclidean searches (matchFeaturesScratch.m:142-155) against the oracle's restatement, through
    featureMatchingPairwise and matchFeaturesScratch; binary descriptors take the exhaustive path; 'pca2nn' raises."""
    import warnings

    ctx = aps._lib.default_context()
    inp = {"Matchingmethod": "Approximate", "ApproxFloatNNMethod": nn, "Matchingthreshold": 1.5, "Ratiothreshold": 0.7,
           "useMATLABFeatureMatch": 0}
    for cid, n, kp in ((5, 8, 2500), (1, 5, 1500)):
        desc, _ = aps.synth.make_config(cid, n=n, kp=kp)
        with warnings.catch_warnings():
            warnings.simplefilter("error")                       # built modes must not warn
            got, met = aps.featureMatchingPairwise(inp, desc, len(desc), ctx=ctx, return_metric=True)
        total = 0
        for j in range(len(desc)):
            for i in range(j):
                if desc[i].shape[0] == 0 or desc[j].shape[0] == 0:
                    assert got[i][j].shape[0] == 0
                    continue
                m, d = orc.match_features_method(desc[i], desc[j], 1.5, 0.7, nn)
                assert got[i][j].shape == (len(m), 2) and np.array_equal(got[i][j], m.astype(np.float64)), (cid, i, j)
                if len(m):
                    assert np.array_equal(np.asarray(met[i][j]), d), (cid, i, j)
                total += len(m)
        assert total > 500
    A, B = desc[0], desc[1]
    m, d = aps.matchFeaturesScratch(A, B, Method="Approximate", ApproxFloatNNMethod=nn, MatchThreshold=1.5, MaxRatio=0.7, ctx=ctx)
    om, od = orc.match_features_method(A, B, 1.5, 0.7, nn)
    assert np.array_equal(m, om) and np.array_equal(d, od)
    # binary: exhaustive whatever the method says (matchFeaturesScratch.m:611)
    orb, _ = aps.synth.make_config(4, n=3, kp=800)
    gb = aps.featureMatchingPairwise(dict(inp, Matchingthreshold=40.0), [aps.binaryFeatures(x) for x in orb], 3, ctx=ctx)
    for j in range(3):
        for i in range(j):
            m, _ = orc.match_features(orb[i], orb[j], 40.0, 0.7)
            assert np.array_equal(gb[i][j], m.astype(np.float64))
    with pytest.raises(ValueError):
        aps.featureMatchingPairwise(dict(inp, ApproxFloatNNMethod="lsh"), desc, len(desc), ctx=ctx)


def test_subsetpdist2_above_the_subset_size(aps, orc):
    """'subsetpdist2' when a train image has more rows than the subset (12000 in the reference, matchFeaturesScratch.m:151;
    small numbers here): candB is `subset` DISTINCT rows in a pseudo-random order, the same for every pair with that
    train image; the match lists equal the oracle's run on the same candB; images at or below the size use all rows."""
    ctx = aps._lib.default_context()
    desc, _ = aps.synth.make_config(5, n=6, kp=3000)
    desc = [d[:c] for d, c in zip(desc, (3000, 900, 2600, 3000, 1500, 2999))]
    subset = 1500
    plan = aps.PairwisePlan(ctx, [d.shape[0] for d in desc], 64, False)
    plan.set_method("subsetpdist2", subset=subset, seed=7)
    plan.upload(desc)
    plan.prepare()
    pp, rows, met = plan.match(1.5, 0.75)
    tables = {}
    for j, d in enumerate(desc):
        if d.shape[0] > subset:
            t = plan.subset_table(j)
            assert t.shape == (subset,) and len(set(t.tolist())) == subset and t.min() >= 0 and t.max() < d.shape[0]
            assert not np.array_equal(t, np.arange(subset))          # a permutation prefix, not the first rows
            tables[j] = t
        else:
            with pytest.raises(aps.ApsError):
                plan.subset_table(j)
    assert sorted(tables) == [0, 2, 3, 5]
    plan2 = aps.PairwisePlan(ctx, [d.shape[0] for d in desc], 64, False)     # another seed, another subset
    plan2.set_method("subsetpdist2", subset=subset, seed=8)
    plan2.upload(desc)
    plan2.prepare()
    assert not np.array_equal(plan2.subset_table(0), tables[0])
    plan2.close()
    n, total = len(desc), 0
    for j in range(n):
        for i in range(j):
            c = i + j * n
            g, gm = rows[pp[c]:pp[c + 1]], met[pp[c]:pp[c + 1]]
            if j in tables:
                m, d = orc.match_features_subset(desc[i], desc[j], tables[j], 1.5, 0.75)
            else:
                m, d = orc.match_features_method(desc[i], desc[j], 1.5, 0.75, "subsetpdist2")
            assert np.array_equal(g, m) and np.array_equal(gm, d), (i, j)
            total += len(m)
    assert total > 300
    plan.close()


def test_pca2nn_matches_the_oracle(aps, orc):
    """ApproxFloatNNMethod 'pca2nn' (matchFeaturesScratch.m:442-573): PCA basis of the train image (mean, float64
    covariance, cyclic Jacobi -- csrc/aps_pca.cu repeats the oracle's arithmetic operation by operation), both images
    projected on 48 components, rows re-normalised, cosine 2-NN, d = 2 - 2 sim.  Match lists and metrics equal the
    oracle's bit for bit; D <= 48 skips the PCA (:477)."""
    ctx = aps._lib.default_context()
    inp = {"Matchingmethod": "Approximate", "ApproxFloatNNMethod": "pca2nn", "Matchingthreshold": 1.5, "Ratiothreshold": 0.7,
           "useMATLABFeatureMatch": 0}
    sets = [aps.synth.make_config(5, n=6, kp=1500)[0],               # KAZE-64
            aps.synth.make_config(1, n=4, kp=1200)[0],               # integer SIFT-128: front end normalises first
            [d[:, :40].copy() for d in aps.synth.make_config(5, n=4, kp=900)[0]]]   # D = 40: no PCA
    for si, desc in enumerate(sets):
        for screen in (True, False):
            ctx.set_pairwise_screen(screen)
            try:
                got, met = aps.featureMatchingPairwise(inp, desc, len(desc), ctx=ctx, return_metric=True)
            finally:
                ctx.set_pairwise_screen(True)
            total = 0
            for j in range(len(desc)):
                for i in range(j):
                    if desc[i].shape[0] == 0 or desc[j].shape[0] == 0:
                        assert got[i][j].shape[0] == 0
                        continue
                    m, d = orc.match_features_pca(desc[i], desc[j], 1.5, 0.7)
                    assert got[i][j].shape == (len(m), 2) and np.array_equal(got[i][j], m.astype(np.float64)), (si, screen, i, j)
                    if len(m):
                        assert np.array_equal(np.asarray(met[i][j]), d), (si, screen, i, j)
                    total += len(m)
            assert total > 200, (si, total)
    A, B = sets[0][0], sets[0][1]
    m, d = aps.matchFeaturesScratch(A, B, Method="Approximate", MatchThreshold=1.5, MaxRatio=0.7, ctx=ctx)   # parser default = pca2nn
    om, od = orc.match_features_pca(A, B, 1.5, 0.7)
    assert np.array_equal(m, om) and np.array_equal(d, od)


@pytest.mark.gpu
@pytest.mark.parametrize("D", [30, 61, 36])
def test_descriptor_lengths_not_a_multiple_of_four(aps, orc, D):
    """The sequential exact-distance helpers (aps_exact_math.cuh) read rows with 128-bit loads only when the row length
    is a multiple of four and the rows are 16-byte aligned; D = 30 / 61 take the scalar loop, D = 36 the vector loop with a
    padded operand length (Dp = 64).  Exhaustive and Euclidean ('subsetpdist2') modes against the oracle, matches and
    metric bit for bit."""
    rng = np.random.default_rng(900 + D)
    A = rng.standard_normal((1500, D)).astype(np.float32)
    B = rng.standard_normal((1700, D)).astype(np.float32)
    B[:400] = A[200:600] + 0.01 * rng.standard_normal((400, D)).astype(np.float32)
    B[400:420] = A[700:720]                                   # exact duplicates: ties
    m, met = aps.matchFeaturesScratch(A, B, MatchThreshold=1.5, MaxRatio=0.7)
    om, omet = orc.match_features(A, B, 1.5, 0.7)
    assert len(om) > 300 and np.array_equal(m, om) and np.array_equal(met, omet), D
    m, met = aps.matchFeaturesScratch(A, B, Method="Approximate", ApproxFloatNNMethod="subsetpdist2", MatchThreshold=1.5,
                                      MaxRatio=0.7)
    om, omet = orc.match_features_method(A, B, 1.5, 0.7, "subsetpdist2")
    assert len(om) > 300 and np.array_equal(m, om) and np.array_equal(met, omet), D
