"""ctypes bindings for the CPU oracle (oracle/liborc.so) and the verbatim reference build
(oracle/_ref/*.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.

Every function mirrors one C entry of aps_oracle.c (which cites the reference file:line).
Matrices are numpy row-major; indices returned are 1-based like the reference's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")
_REF_DIR = os.path.join(_HERE, "_ref")

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile liborc.so (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    if ref and os.path.isdir("/root/reference/Procedural Program/mex"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build(ref=False)
        L = C.CDLL(_LIB)
        L.orc_num_threads.restype = C.c_int
        L.orc_normalize_rows_global.argtypes = [_f32p, C.c_int64, C.c_int]
        L.orc_normalize_rows_pairwise.argtypes = [_f32p, C.c_int64, C.c_int]
        L.orc_knn_l2.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_int, _u32p, _f32p]
        L.orc_knn_hamming.argtypes = [_u8p, C.c_int64, _u8p, C.c_int64, C.c_int, C.c_int, _u32p, _f32p]
        L.orc_nearest2_hamming.argtypes = [_u8p, C.c_int64, _u8p, C.c_int64, C.c_int, _u32p, _f32p, _f32p]
        L.orc_nearest2_ssd.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, _u32p, _f32p, _f32p]
        L.orc_global_filter.argtypes = [_u32p, _f32p, C.c_int64, C.c_int, _i64p, C.c_int, C.c_double, C.c_int,
                                        _i32p, _u32p, C.POINTER(C.c_int64)]
        L.orc_global_scatter.argtypes = [_i32p, _u32p, C.c_int64, _i64p, C.c_int, _i64p, _u32p]
        L.orc_global_scatter.restype = C.c_int64
        L.orc_feature_matching_global.argtypes = [C.c_void_p, C.c_int, _i64p, C.c_int, C.c_int, C.c_int,
                                                  C.c_double, C.c_int, _i64p, _u32p, C.c_void_p, C.c_void_p,
                                                  C.POINTER(C.c_int64)]
        L.orc_feature_matching_global.restype = C.c_int64
        L.orc_filter_unique.argtypes = [_u32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                        C.c_double, C.c_int, _u32p, _f64p]
        L.orc_filter_unique.restype = C.c_int64
        L.orc_match_features.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                         C.c_double, C.c_double, C.c_int, _u32p, _f64p]
        L.orc_match_features.restype = C.c_int64
        L.orc_match_features_method.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                                C.c_double, C.c_double, C.c_int, _u32p, _f64p]
        L.orc_match_features_method.restype = C.c_int64
        L.orc_match_features_subset.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, _i32p, C.c_int64, C.c_double,
                                                C.c_double, C.c_int, _u32p, _f64p]
        L.orc_match_features_subset.restype = C.c_int64
        L.orc_match_features_pca.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int,
                                             _u32p, _f64p]
        L.orc_match_features_pca.restype = C.c_int64
        L.orc_nearest2_euclid.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, _u32p, _f32p, _f32p]
        L.orc_feature_matching_pairwise.argtypes = [C.c_void_p, C.c_int, _i64p, C.c_int, C.c_int, C.c_double,
                                                    C.c_double, _i64p, _u32p, _f64p]
        L.orc_feature_matching_pairwise.restype = C.c_int64
        L.orc_select_partners.argtypes = [_i64p, C.c_int, C.c_int, _u8p, _i64p]
        L.orc_select_partners.restype = C.c_int64
        L.orc_pack_bits.argtypes = [_u8p, C.c_int64, C.c_int, _u8p]
        L.orc_l2sq.argtypes = [_f32p, _f32p, C.c_int]
        L.orc_l2sq.restype = C.c_float
        L.orc_ransac_homography.argtypes = [_f64p, _f64p, C.c_int64, C.c_double, C.c_double, C.c_int, _u32p, C.c_int64,
                                            _f64p, _u8p, _i32p, _i32p]
        L.orc_ransac_homography.restype = C.c_int
        L.orc_inv3.argtypes = [_f64p, _f64p]
        L.orc_image_matching_batch.argtypes = [C.c_int64, _i64p, _f64p, _f64p, C.c_double, C.c_double, C.c_int, _u32p,
                                               C.c_int64, _f64p, _f64p, _u8p, _i32p, _u8p, _i32p]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle's parallel loops (bench.py: all host cores, whatever OMP_NUM_THREADS says)."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def normalize_rows_global(X):
    X = _f32(X).copy()
    lib().orc_normalize_rows_global(X, X.shape[0], X.shape[1])
    return X


def normalize_rows_pairwise(X):
    X = _f32(X).copy()
    lib().orc_normalize_rows_pairwise(X, X.shape[0], X.shape[1])
    return X


def knn_l2(train, query, k):
    train, query = _f32(train), _f32(query)
    Fq = query.shape[0]
    idx = np.zeros((Fq, k), np.uint32)
    dist = np.zeros((Fq, k), np.float32)
    lib().orc_knn_l2(train, train.shape[0], query, Fq, train.shape[1], k, idx, dist)
    return idx, dist


def knn_hamming(train, query, k):
    train, query = _u8(train), _u8(query)
    Fq = query.shape[0]
    idx = np.zeros((Fq, k), np.uint32)
    dist = np.zeros((Fq, k), np.float32)
    lib().orc_knn_hamming(train, train.shape[0], query, Fq, train.shape[1], k, idx, dist)
    return idx, dist


def nearest2_hamming(A, B):
    A, B = _u8(A), _u8(B)
    N1 = A.shape[0]
    idx2 = np.zeros(N1, np.uint32)
    d1 = np.zeros(N1, np.float32)
    d2 = np.zeros(N1, np.float32)
    lib().orc_nearest2_hamming(A, N1, B, B.shape[0], A.shape[1], idx2, d1, d2)
    return idx2, d1, d2


def nearest2_ssd(A, B):
    A, B = _f32(A), _f32(B)
    N1 = A.shape[0]
    idx2 = np.zeros(N1, np.uint32)
    d1 = np.zeros(N1, np.float32)
    d2 = np.zeros(N1, np.float32)
    lib().orc_nearest2_ssd(A, N1, B, B.shape[0], A.shape[1], idx2, d1, d2)
    return idx2, d1, d2


def global_filter(idx, dist, counts, ratio, ratio_mode=0):
    idx = np.ascontiguousarray(idx, np.uint32)
    dist = _f32(dist)
    counts = np.ascontiguousarray(counts, np.int64)
    F, k = idx.shape
    tgt = np.zeros(F, np.int32)
    par = np.zeros(F, np.uint32)
    amb = C.c_int64(0)
    lib().orc_global_filter(idx, dist, F, k, counts, len(counts), float(ratio), ratio_mode, tgt, par, C.byref(amb))
    return tgt, par, int(amb.value)


def _csr_to_cells(pair_ptr, rows, n):
    """CSR -> dict {(i,j) 0-based: [M x 2] uint32}; only non-empty cells are present."""
    cells = {}
    for j in range(n):
        for i in range(n):
            c = i + j * n
            a, b = int(pair_ptr[c]), int(pair_ptr[c + 1])
            if b > a:
                cells[(i, j)] = rows[a:b].copy()
    return cells


def feature_matching_global(desc_list, k, ratio, ratio_mode=0, return_knn=False):
    """featureMatchingGlobal with an exact kNN.  desc_list: list of [Ni x D] float32 or uint8."""
    n = len(desc_list)
    counts = np.array([d.shape[0] for d in desc_list], np.int64)
    F = int(counts.sum())
    nonempty = [d for d in desc_list if d.shape[0] > 0]
    is_binary = bool(nonempty) and nonempty[0].dtype == np.uint8
    D = nonempty[0].shape[1] if nonempty else 0
    pair_ptr = np.zeros(n * n + 1, np.int64)
    rows = np.zeros((max(F, 1), 2), np.uint32)
    amb = C.c_int64(0)
    knn_idx = np.zeros((max(F, 1), k), np.uint32)
    knn_dist = np.zeros((max(F, 1), k), np.float32)
    if F > 0:
        pooled = np.ascontiguousarray(np.concatenate(nonempty, axis=0))
        M = lib().orc_feature_matching_global(pooled.ctypes.data, int(is_binary), counts, n, D, k, float(ratio),
                                              ratio_mode, pair_ptr, rows.reshape(-1), knn_idx.ctypes.data,
                                              knn_dist.ctypes.data, C.byref(amb))
    else:
        M = 0
    out = {"pair_ptr": pair_ptr, "rows": rows[:M], "cells": _csr_to_cells(pair_ptr, rows, n),
           "n_ambiguous": int(amb.value)}
    if return_knn:
        out["knn_idx"], out["knn_dist"] = knn_idx[:F], knn_dist[:F]
    return out


def match_features(A, B, match_threshold, max_ratio, unique=True):
    """matchFeaturesScratch(A,B,'Method','Exhaustive','MatchThreshold',..,'MaxRatio',..,'Unique',..)."""
    is_binary = A.dtype == np.uint8
    A = _u8(A) if is_binary else _f32(A)
    B = _u8(B) if is_binary else _f32(B)
    N1, N2 = A.shape[0], B.shape[0]
    m = np.zeros((max(N1, 1), 2), np.uint32)
    met = np.zeros(max(N1, 1), np.float64)
    K = lib().orc_match_features(A.ctypes.data, N1, B.ctypes.data, N2, A.shape[1], int(is_binary),
                                 float(match_threshold), float(max_ratio), int(unique), m.reshape(-1), met)
    return m[:K].copy(), met[:K].copy()


METHODS = {"exhaustive": 0, "subsetpdist2": 1, "kdtree": 2}


def match_features_method(A, B, match_threshold, max_ratio, method, unique=True):
    """matchFeaturesScratch(A,B,'Method','Approximate','ApproxFloatNNMethod',method,...) -- 'subsetpdist2' (while
    N2 <= 12000) and 'kdtree' are exact Euclidean searches; method 'exhaustive' = match_features."""
    is_binary = A.dtype == np.uint8
    A = _u8(A) if is_binary else _f32(A)
    B = _u8(B) if is_binary else _f32(B)
    N1, N2 = A.shape[0], B.shape[0]
    m = np.zeros((max(N1, 1), 2), np.uint32)
    met = np.zeros(max(N1, 1), np.float64)
    K = lib().orc_match_features_method(A.ctypes.data, N1, B.ctypes.data, N2, A.shape[1], int(is_binary),
                                        METHODS[method], float(match_threshold), float(max_ratio), int(unique),
                                        m.reshape(-1), met)
    return m[:K].copy(), met[:K].copy()


def match_features_pca(A, B, match_threshold, max_ratio, unique=True):
    """matchFeaturesScratch(A,B,'Method','Approximate','ApproxFloatNNMethod','pca2nn',...) for float descriptors."""
    A, B = _f32(A), _f32(B)
    N1 = A.shape[0]
    m = np.zeros((max(N1, 1), 2), np.uint32)
    met = np.zeros(max(N1, 1), np.float64)
    K = lib().orc_match_features_pca(A, N1, B, B.shape[0], A.shape[1], float(match_threshold), float(max_ratio),
                                     int(unique), m.reshape(-1), met)
    return m[:K].copy(), met[:K].copy()


def match_features_subset(A, B, candB, match_threshold, max_ratio, unique=True):
    """'subsetpdist2' with more rows in B than the subset: candB = 0-based candidate rows in randperm order."""
    A, B = _f32(A), _f32(B)
    candB = np.ascontiguousarray(candB, np.int32)
    N1 = A.shape[0]
    m = np.zeros((max(N1, 1), 2), np.uint32)
    met = np.zeros(max(N1, 1), np.float64)
    K = lib().orc_match_features_subset(A, N1, B, B.shape[0], A.shape[1], candB, candB.size, float(match_threshold),
                                        float(max_ratio), int(unique), m.reshape(-1), met)
    return m[:K].copy(), met[:K].copy()


def nearest2_euclid(A, B):
    A, B = _f32(A), _f32(B)
    N1 = A.shape[0]
    idx2, d1, d2 = np.zeros(N1, np.uint32), np.zeros(N1, np.float32), np.zeros(N1, np.float32)
    lib().orc_nearest2_euclid(A, N1, B, B.shape[0], A.shape[1], idx2, d1, d2)
    return idx2, d1, d2


def feature_matching_pairwise(desc_list, match_threshold, max_ratio):
    n = len(desc_list)
    counts = np.array([d.shape[0] for d in desc_list], np.int64)
    nonempty = [d for d in desc_list if d.shape[0] > 0]
    is_binary = bool(nonempty) and nonempty[0].dtype == np.uint8
    D = nonempty[0].shape[1] if nonempty else 1
    cap = int(sum(int(counts[i]) * (n - 1 - i) for i in range(n)))
    pair_ptr = np.zeros(n * n + 1, np.int64)
    rows = np.zeros((max(cap, 1), 2), np.uint32)
    metric = np.zeros(max(cap, 1), np.float64)
    if nonempty:
        pooled = np.ascontiguousarray(np.concatenate(nonempty, axis=0))
        M = lib().orc_feature_matching_pairwise(pooled.ctypes.data, int(is_binary), counts, n, D,
                                                float(match_threshold), float(max_ratio), pair_ptr,
                                                rows.reshape(-1), metric)
    else:
        M = 0
    return {"pair_ptr": pair_ptr, "rows": rows[:M], "metric": metric[:M],
            "cells": _csr_to_cells(pair_ptr, rows, n)}


def select_partners(counts_nn, m):
    """imageMatching.m:75-100.  counts_nn: [n x n] array, counts_nn[i, j] = rows(matches{i+1,j+1})."""
    cm = np.ascontiguousarray(np.asarray(counts_nn, np.int64).T)  # column-major buffer
    n = cm.shape[0]
    cand = np.zeros(n * n, np.uint8)
    pairs = np.zeros(max(n * n, 1), np.int64)
    npairs = lib().orc_select_partners(cm.reshape(-1), n, int(m), cand, pairs)
    cand_ij = cand.reshape(n, n).T.astype(bool)  # back to [i, j]
    return cand_ij, pairs[:npairs].copy()


def pack_bits(bits01):
    bits01 = _u8(bits01)
    N, Db = bits01.shape
    out = np.zeros((N, (Db + 7) // 8), np.uint8)
    lib().orc_pack_bits(bits01, N, Db, out)
    return out


# --------------------------------------------------------------------------------------------
# consumer of the match lists: RANSAC homographies (aps_oracle_ransac.c)
# --------------------------------------------------------------------------------------------
def ransac_homography(pts1, pts2, max_distance, confidence, max_trials, samples):
    """estimateTransformationRANSAC.m:94-183 ('projective').  pts1/pts2: [n x 2] float64 (matchedPoints1/2),
    samples: [n_draws x 4] zero-based minimal samples.  -> (found, model[3x3] row-major, inliers bool[n], draws_used)"""
    p1 = np.ascontiguousarray(pts1, np.float64).reshape(-1, 2)
    p2 = np.ascontiguousarray(pts2, np.float64).reshape(-1, 2)
    smp = np.ascontiguousarray(samples, np.uint32).reshape(-1, 4)
    n = p1.shape[0]
    model = np.zeros(9, np.float64)
    inl = np.zeros(max(n, 1), np.uint8)
    ni, du = np.zeros(1, np.int32), np.zeros(1, np.int32)
    found = lib().orc_ransac_homography(p1.reshape(-1), p2.reshape(-1), n, float(max_distance), float(confidence),
                                        int(max_trials), smp.reshape(-1), smp.shape[0], model, inl, ni, du)
    return bool(found), model.reshape(3, 3), inl[:n].astype(bool), int(du[0])


def image_matching_batch(pt_ptr, pts1, pts2, max_distance, confidence, max_trials, samples):
    """imageMatching.m:121-156 for a batch of candidate pairs (CSR correspondences).  samples: [n_pairs x n_draws x 4].
    -> dict(models [P x 3 x 3], models_inv, inliers bool[total], n_inliers, accepted, draws_used)"""
    pt_ptr = np.ascontiguousarray(pt_ptr, np.int64)
    P = pt_ptr.size - 1
    p1 = np.ascontiguousarray(pts1, np.float64).reshape(-1, 2)
    p2 = np.ascontiguousarray(pts2, np.float64).reshape(-1, 2)
    smp = np.ascontiguousarray(samples, np.uint32).reshape(P, -1, 4)
    total = p1.shape[0]
    models, minv = np.zeros((P, 9)), np.zeros((P, 9))
    inl = np.zeros(max(total, 1), np.uint8)
    ni, du, acc = np.zeros(max(P, 1), np.int32), np.zeros(max(P, 1), np.int32), np.zeros(max(P, 1), np.uint8)
    lib().orc_image_matching_batch(P, pt_ptr, p1.reshape(-1), p2.reshape(-1), float(max_distance), float(confidence),
                                   int(max_trials), smp.reshape(-1), smp.shape[1], models.reshape(-1), minv.reshape(-1),
                                   inl, ni, acc, du)
    return dict(models=models.reshape(P, 3, 3), models_inv=minv.reshape(P, 3, 3), inliers=inl[:total].astype(bool),
                n_inliers=ni[:P], accepted=acc[:P].astype(bool), draws_used=du[:P])


# --------------------------------------------------------------------------------------------
# verbatim reference build (oracle/_ref)
# --------------------------------------------------------------------------------------------
_ref_libs = {}


def ref_available(name="nearest2HammingExhaustiveMEX") -> bool:
    return os.path.exists(os.path.join(_REF_DIR, name + ".so"))


def ref_nearest2_hamming(A, B, omp=False):
    """Runs the reference's own nearest2HammingExhaustive{,OMP}MEX.cpp (compiled verbatim)."""
    name = "nearest2HammingExhaustiveOMPMEX" if omp else "nearest2HammingExhaustiveMEX"
    if name not in _ref_libs:
        L = C.CDLL(os.path.join(_REF_DIR, name + ".so"))
        L.ref_nearest2_hamming.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, _u32p, _f32p,
                                           _f32p, C.c_char_p, C.c_int]
        L.ref_nearest2_hamming.restype = C.c_int
        _ref_libs[name] = L
    L = _ref_libs[name]
    A, B = _u8(A), _u8(B)
    N1, N2, nb = A.shape[0], B.shape[0], A.shape[1]
    Acm = np.asfortranarray(A)  # the MEX boundary is column-major
    Bcm = np.asfortranarray(B)
    idx2 = np.zeros(max(N1, 1), np.uint32)
    d1 = np.zeros(max(N1, 1), np.float32)
    d2 = np.zeros(max(N1, 1), np.float32)
    err = C.create_string_buffer(512)
    rc = L.ref_nearest2_hamming(Acm.ctypes.data, N1, Bcm.ctypes.data, N2, nb, idx2, d1, d2, err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return idx2[:N1], d1[:N1], d2[:N1]
