"""Mints tests/golden/matlab_semantics_v1.npz from the numpy restatement of the MATLAB-only pieces
(tests/matlab_restatement.py) ALONE -- neither the C oracle nor the CUDA library is involved.  The oracle is
checked against the file by tests/test_oracle_matlab_restatement.py, the GPU path by tests/test_gpu_parity.py.
Run from the repo root:  python tests/golden/make_golden_semantics.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import matlab_restatement as mr  # noqa: E402

rng = np.random.default_rng(20071017)
G = {}

# matchFeaturesScratch, float: |x| <= 2 (no normalisation), small integers => exact SSD, plenty of ties
A = rng.integers(-2, 3, (180, 24)).astype(np.float32)
B = A[rng.permutation(180)[:140]].copy()
B[rng.random(B.shape) < 0.02] += 1.0
B = np.clip(B, -2, 2)
G["mf_A"], G["mf_B"], G["mf_thr"], G["mf_ratio"] = A, B, np.float64(3.5), np.float64(0.8)
m, met = mr.match_features(A, B, 3.5, 0.8, True)
G["mf_matches"], G["mf_metric"] = m, met.astype(np.float64)

# matchFeaturesScratch, packed binary (32-bit descriptors: Hamming ties)
A = rng.integers(0, 256, (150, 4), dtype=np.uint8)
B = A[rng.permutation(150)[:120]] ^ (rng.random((120, 4)) < 0.1).astype(np.uint8)
G["mb_A"], G["mb_B"], G["mb_thr"], G["mb_ratio"] = A, B, np.float64(25.0), np.float64(0.8)
m, met = mr.match_features(A, B, 25.0, 0.8, True)
G["mb_matches"], G["mb_metric"] = m, met.astype(np.float64)

# featureMatchingGlobal, binary descriptors (exact integer distances), duplicates inside and across images, empty image
counts = np.array([35, 0, 50, 28, 41], np.int64)
base = rng.integers(0, 256, (70, 8), dtype=np.uint8)
desc = np.concatenate([base[rng.integers(0, 70, c)] ^ (rng.random((c, 8)) < 0.12).astype(np.uint8) for c in counts])
G["gl_desc"], G["gl_counts"], G["gl_k"], G["gl_ratio"] = desc, counts, np.int64(4), np.float64(0.8)
pp, rows = mr.feature_matching_global(np.split(desc, np.cumsum(counts)[:-1]), 4, 0.8)
G["gl_pair_ptr"], G["gl_rows"] = pp, rows

# featureMatchingGlobal, float descriptors (real-valued: no near-ties), planted near-copies across images
counts = np.array([60, 45, 0, 52], np.int64)
src = rng.standard_normal((80, 32)).astype(np.float32)
parts = []
for c in counts:
    pick = rng.integers(0, 80, c)
    parts.append((src[pick] + rng.normal(0, 0.05, (c, 32))).astype(np.float32))
desc = np.concatenate(parts)
desc[70] = desc[3]                                   # exact duplicate across images: self need not come first
G["gf_desc"], G["gf_counts"], G["gf_k"], G["gf_ratio"] = desc, counts, np.int64(4), np.float64(0.8)
pp, rows = mr.feature_matching_global(np.split(desc, np.cumsum(counts)[:-1]), 4, 0.8)
G["gf_pair_ptr"], G["gf_rows"] = pp, rows

# imageMatching top-m partners: small counts => ties, zero rows
C = np.triu(rng.integers(0, 4, (12, 12)), 1)
cand, lin = mr.select_partners(C, 4)
G["sp_counts"], G["sp_m"], G["sp_cand"], G["sp_lin"] = C, np.int64(4), cand, lin

out = os.path.join(HERE, "matlab_semantics_v1.npz")
np.savez_compressed(out, **G)
print("wrote", out, os.path.getsize(out), "bytes;", {k: (v.shape if hasattr(v, "shape") else v) for k, v in G.items() if k.endswith(("matches", "rows", "lin"))})
