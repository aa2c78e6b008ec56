"""Generates tests/golden/golden_v1.npz -- run ONLY in the build container (needs /root/reference
for oracle/_ref and cv2 for the OpenCV engines).  The .npz is committed; tests never regenerate it.

Sources of truth recorded here:
  ham2nn_*  : outputs of the reference's own nearest2HammingExhaustive{,OMP}MEX.cpp compiled verbatim
              (oracle/_ref; PP/mex/nearest2HammingExhaustiveMEX.cpp:16-80, ...OMPMEX.cpp:18-83)
  bfknn_*   : cv2.BFMatcher(NORM_HAMMING, False).knnMatch -- the engine behind
              flann_knn_win(...,'bf') (PP/mex/flann_knn.cpp:199-223); cv2 4.13 here, 4.12 pinned there
  flann_*   : cv2.flann_Index(algorithm=LINEAR).knnSearch -- FLANN's own L2 functor and result
              ordering with an exhaustive search (flann_knn.cpp:229-234 uses the same functor behind
              a randomised KD-tree); squared distances, 0-based indices stored +1
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

rng = np.random.Generator(np.random.PCG64(20070099))
G = {}


def ham_case(name, A, B):
    G[f"ham2nn_{name}_A"], G[f"ham2nn_{name}_B"] = A, B
    i, d1, d2 = oracle.ref_nearest2_hamming(A, B, omp=False)
    io, d1o, d2o = oracle.ref_nearest2_hamming(A, B, omp=True)
    assert (i == io).all() and np.array_equal(d1, d1o, equal_nan=True) and np.array_equal(d2, d2o, equal_nan=True)
    G[f"ham2nn_{name}_idx2"], G[f"ham2nn_{name}_d1"], G[f"ham2nn_{name}_d2"] = i, d1, d2


ham_case("rand256", rng.integers(0, 256, (64, 32), dtype=np.uint8), rng.integers(0, 256, (96, 32), dtype=np.uint8))
ham_case("ties", rng.integers(0, 4, (50, 2), dtype=np.uint8), rng.integers(0, 4, (70, 2), dtype=np.uint8))
ham_case("n2is1", rng.integers(0, 256, (5, 32), dtype=np.uint8), rng.integers(0, 256, (1, 32), dtype=np.uint8))
ham_case("n2is0", rng.integers(0, 256, (5, 32), dtype=np.uint8), np.zeros((0, 32), np.uint8))
B = rng.integers(0, 256, (40, 64), dtype=np.uint8)
A = B[rng.permutation(40)[:30]].copy()  # exact duplicates -> d1 == 0, BRISK-width rows
ham_case("dups512", A, B)


def bf_case(name, Q, T, k):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(Q, T, k)
    idx = np.zeros((len(Q), k), np.uint32)
    dist = np.full((len(Q), k), np.inf, np.float32)
    for r, row in enumerate(m):
        for c, x in enumerate(row):
            idx[r, c], dist[r, c] = x.trainIdx + 1, x.distance
    G[f"bfknn_{name}_Q"], G[f"bfknn_{name}_T"] = Q, T
    G[f"bfknn_{name}_idx"], G[f"bfknn_{name}_dist"] = idx, dist


bf_case("rand256", rng.integers(0, 256, (80, 32), dtype=np.uint8), rng.integers(0, 256, (120, 32), dtype=np.uint8), 4)
bf_case("ties", rng.integers(0, 4, (60, 2), dtype=np.uint8), rng.integers(0, 4, (90, 2), dtype=np.uint8), 4)
T = rng.integers(0, 256, (100, 32), dtype=np.uint8)
bf_case("self", T, T, 4)
bf_case("kgtF", rng.integers(0, 256, (6, 32), dtype=np.uint8), rng.integers(0, 256, (3, 32), dtype=np.uint8), 4)


def flann_case(name, Q, T, k):
    i, d = cv2.flann_Index(T, dict(algorithm=0)).knnSearch(Q, k, params=dict(checks=32))
    G[f"flann_{name}_Q"], G[f"flann_{name}_T"] = Q, T
    G[f"flann_{name}_idx"], G[f"flann_{name}_dist"] = (i + 1).astype(np.uint32), d.astype(np.float32)


def sift_like(n):
    v = np.abs(rng.standard_normal((n, 128))).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = np.minimum(v, 0.2)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return np.minimum(np.rint(512 * v), 255).astype(np.float32)


S = sift_like(160)
S[150:160] = S[10:20]  # exact duplicates: self is not the unique zero-distance neighbour
Sn = oracle.normalize_rows_global(S)
flann_case("sift_self", Sn, Sn, 4)
K = rng.standard_normal((140, 64)).astype(np.float32)
K /= np.linalg.norm(K, axis=1, keepdims=True)
flann_case("kaze", K[:60].copy(), K[40:].copy(), 4)
R = (rng.standard_normal((50, 37)) * 3).astype(np.float32)  # D not a multiple of 4: scalar tail of the functor
flann_case("tail37", R[:20].copy(), R, 3)

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
np.savez_compressed(out, **G)
print("wrote", out, os.path.getsize(out), "bytes,", len(G), "arrays; cv2", cv2.__version__)
