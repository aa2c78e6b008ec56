"""Generates tests/golden/ransac_v1.npz.  The .npz is committed; tests never regenerate it.

Source of truth: tests/ransac_lapack.py -- estimateTransformationRANSAC.m ('projective') restated with
numpy.linalg (LAPACK svd / solve / cond, the library MATLAB calls for svd, mldivide and rcond).  Neither the
oracle nor the GPU code is involved in producing these vectors.  Each case stores the correspondences, the
table of minimal samples, the parameters and the LAPACK result (found, model, inlier mask, draws consumed)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ransac_lapack as rl  # noqa: E402

G = {}
cases = [("clean80", 120, 0.8, 0.5, 5.5, 99.9, 500, 11), ("half", 300, 0.5, 0.7, 5.5, 99.9, 500, 12),
         ("sparse", 600, 0.3, 0.6, 3.0, 99.0, 800, 13), ("tight", 200, 0.7, 0.2, 1.5, 99.9, 300, 14),
         ("tiny", 9, 1.0, 0.1, 5.5, 99.9, 500, 15), ("big", 2000, 0.6, 0.8, 5.5, 99.9, 500, 16)]
for name, n, frac, noise, md, conf, mt, seed in cases:
    rng = np.random.default_rng(seed)
    p1, p2, H = rl.make_pair(rng, n, frac, noise=noise)
    smp = rl.draw_table(rng, n, 2 * mt)
    found, model, inl, used = rl.ransac(p1, p2, md, conf, mt, smp)
    assert found and inl.sum() > 8
    G[f"{name}_p1"], G[f"{name}_p2"], G[f"{name}_samples"] = p1, p2, smp
    G[f"{name}_params"] = np.array([md, conf, mt], np.float64)
    G[f"{name}_model"], G[f"{name}_inliers"], G[f"{name}_draws"] = model, inl, np.int32(used)
    # margin of the inlier decision: smallest |error - threshold| under the final model (how far from a tie)
    ph1, ph2 = np.c_[p1, np.ones(n)], np.c_[p2, np.ones(n)]
    _, err = rl.find_inliers(model, ph1, ph2, md)
    G[f"{name}_margin"] = np.float64(np.min(np.abs(err[np.isfinite(err)] - md)))
    print(name, n, int(inl.sum()), used, float(G[f"{name}_margin"]))
G["cases"] = np.array([c[0] for c in cases])
np.savez_compressed(os.path.join(HERE, "ransac_v1.npz"), **G)
