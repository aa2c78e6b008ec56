"""Independent numpy restatement of the MATLAB-only pieces of the path -- TEST INFRASTRUCTURE.

MATLAB / Octave are absent from this image and the reference ships no tests, so the C oracle's MATLAB-derived
functions cannot be pinned against the reference itself.  This module is a SECOND restatement, written
array-style straight from the .m text (1-based MATLAB semantics emulated with numpy, no code shared with
oracle/aps_oracle.c), used to cross-check the oracle on random inputs (tests/test_oracle_matlab_restatement.py)
and to mint the golden vectors tests/golden/matlab_semantics_v1.npz (make_golden_semantics.py).

PP/ = /root/reference/Procedural Program/.  Each function cites the lines it follows.
"""
import numpy as np

EPS32 = np.float32(np.finfo(np.float32).eps)


def normalize_global(X):
    """PP/featureMatching/featureMatchingGlobal.m:80-84   single(X) ./ sqrt(sum(X.^2,2) + eps('single'))"""
    X = np.asarray(X, np.float32)
    s = np.zeros(X.shape[0], np.float32)
    for c in range(X.shape[1]):  # column-by-column accumulation in single, like sum(...,2) over a short row
        s = s + X[:, c] * X[:, c]
    return X / np.sqrt(s + EPS32)[:, None]


def normalize_pairwise(X):
    """PP/featureMatching/matchFeaturesScratch.m:232-233   X ./ (sqrt(sum(X.^2,2)) + eps('single'))"""
    X = np.asarray(X, np.float32)
    s = np.zeros(X.shape[0], np.float32)
    for c in range(X.shape[1]):
        s = s + X[:, c] * X[:, c]
    return X / (np.sqrt(s) + EPS32)[:, None]


def global_filter_and_scatter(nn_idx, nn_dist, n_features, ratio_thr):
    """featureMatchingGlobal.m:87-161: imgIdx / localIdx bookkeeping, the per-feature filter loop and the scatter
    into the numImg x numImg cell.  nn_idx is 1-based [F x k] (0 = missing neighbour: dropped here, MATLAB would
    index imgIdx(0) and error), nn_dist [F x k] single.  Returns {(i, j) 1-based: [M x 2] float64}."""
    n_features = [int(v) for v in n_features]
    img_idx = np.repeat(np.arange(1, len(n_features) + 1), n_features)              # :89
    local_idx = np.concatenate([np.arange(1, c + 1) for c in n_features]) if sum(n_features) else np.zeros(0, int)  # :90-97
    cells = {}
    thr = np.float32(ratio_thr)  # mixed single/double comparison: MATLAB compares in single
    for q in range(1, len(img_idx) + 1):                                            # :123
        qi = img_idx[q - 1]
        nidx = nn_idx[q - 1].astype(np.int64)
        ndist = nn_dist[q - 1].astype(np.float32)
        keep = (nidx != q) & (nidx != 0)                                            # :130
        nidx, ndist = nidx[keep], ndist[keep]
        keep = img_idx[nidx - 1] != qi                                              # :135
        nidx, ndist = nidx[keep], ndist[keep]
        if ndist.size < 2:                                                          # :140
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            r = ndist[0] / np.maximum(ndist[1], EPS32)                              # :145 (single / single)
        if r > thr:
            continue
        j = img_idx[nidx[0] - 1]                                                    # :150-152
        li, lj = local_idx[q - 1], local_idx[nidx[0] - 1]
        key, row = ((qi, j), (li, lj)) if qi < j else ((j, qi), (lj, li))           # :155-159
        cells.setdefault(key, []).append(row)
    return {k: np.array(v, np.float64) for k, v in cells.items()}


def nearest2_ssd(A, B):
    """matchFeaturesScratch.m:322-366 with N2 = size(B,1) (the variable is never assigned in the reference):
    blocked a2 + b2' - 2*A*B' in the input class (single), min / mask-with-inf / min.  Returns idx2 (1-based),
    d1, d2 as the reference's double vectors holding single-valued numbers."""
    A, B = np.asarray(A, np.float32), np.asarray(B, np.float32)
    N1, N2 = A.shape[0], B.shape[0]
    block = max(1, int(np.floor(1e7 / max(N2, 1))))                                 # :343
    idx2 = np.zeros(N1, np.uint32)
    d1 = np.full(N1, np.inf)
    d2 = np.full(N1, np.inf)
    for i in range(1, int(np.ceil(N1 / block)) + 1):                                # :346
        s, e = (i - 1) * block + 1, min(N1, i * block)
        Ablk = A[s - 1:e]
        a2 = np.sum(Ablk * Ablk, axis=1, dtype=np.float32)                          # :351
        b2 = np.sum(B * B, axis=1, dtype=np.float32)
        G = Ablk @ B.T                                                              # :353 (sgemm)
        D2 = (a2[:, None] + b2[None, :]) - np.float32(2) * G                        # :354
        idx = np.argmin(D2, axis=1)                                                 # :356 first index on ties
        best = D2[np.arange(D2.shape[0]), idx].copy()
        D2[np.arange(D2.shape[0]), idx] = np.inf                                    # :357
        second = np.min(D2, axis=1) if N2 > 0 else np.full(e - s + 1, np.inf)       # :358
        idx2[s - 1:e] = idx + 1
        d1[s - 1:e] = best
        d2[s - 1:e] = second
    return idx2, d1, d2


def ratio_threshold_unique(idx2, d_best, d_second, n1, n2, is_binary, match_threshold, max_ratio, unique=True):
    """matchFeaturesScratch.m:169-215.  d_best / d_second: percent Hamming (single) or SSD (double vectors).
    Returns (matches [K x 2] uint32 1-based, matchMetric [K])."""
    d_best, d_second = np.asarray(d_best), np.asarray(d_second)
    if is_binary:
        ratio_ok = d_best <= np.float32(max_ratio) * d_second.astype(np.float32)    # :171 single arithmetic
    else:
        r2 = float(max_ratio) * float(max_ratio)                                    # :173
        ratio_ok = d_best.astype(np.float64) <= r2 * d_second.astype(np.float64)
    thresh_ok = d_best <= (np.float32(match_threshold) if is_binary else float(match_threshold))  # :177
    keep = ratio_ok & thresh_ok & np.isfinite(d_best) & np.isfinite(d_second)       # :178
    i1 = np.arange(1, n1 + 1)[keep]                                                 # :180-183
    i2 = np.asarray(idx2, np.int64)[keep]
    d = d_best[keep]
    if unique and i1.size:                                                          # :186
        order = np.argsort(d, kind="stable")                                        # MATLAB sort is stable
        ds, i1s, i2s = d[order], i1[order], i2[order]
        used1 = np.zeros(n1 + 1, bool)
        used2 = np.zeros(n2 + 1, bool)
        kept = np.zeros(i1s.size, bool)
        for t in range(i1s.size):                                                   # :195-204
            a, b = i1s[t], i2s[t]
            if not used1[a] and not used2[b]:
                kept[t] = True
                used1[a] = used2[b] = True
        return np.stack([i1s[kept], i2s[kept]], 1).astype(np.uint32), ds[kept]
    return np.stack([i1, i2], 1).astype(np.uint32), d


def hamming_percent(d1_bits, d2_bits, n_bits):
    """matchFeaturesScratch.m:318 (d2 guard) and :120-121 (percent of mismatched bits, single arithmetic)."""
    d1 = np.asarray(d1_bits, np.float32)
    d2 = np.asarray(d2_bits, np.float32).copy()
    d2[(d2 == 0) | ~np.isfinite(d2)] = np.float32(n_bits)
    return (d1 / np.float32(n_bits)) * np.float32(100), (d2 / np.float32(n_bits)) * np.float32(100)


def nearest2_hamming(A, B):
    """Semantics of PP/mex/nearest2HammingExhaustiveMEX.cpp:38-80 in array form: best = first index attaining
    the minimum, second = second smallest value with multiplicity; N2 == 1 -> second = 8*nb; N2 == 0 -> 0 / NaN."""
    A, B = np.asarray(A, np.uint8), np.asarray(B, np.uint8)
    N1, N2, nb = A.shape[0], B.shape[0], A.shape[1]
    if N2 == 0:
        return np.zeros(N1, np.uint32), np.full(N1, np.nan, np.float32), np.full(N1, np.nan, np.float32)
    H = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(axis=2).astype(np.int64)
    idx = np.argmin(H, axis=1)
    best = H[np.arange(N1), idx]
    if N2 == 1:
        second = np.full(N1, nb * 8, np.int64)
    else:
        H2 = H.copy()
        H2[np.arange(N1), idx] = np.iinfo(np.int64).max
        second = H2.min(axis=1)
    return (idx + 1).astype(np.uint32), best.astype(np.float32), second.astype(np.float32)


def pack_bits(bits01):
    """matchFeaturesScratch.m:617-646: MSB-first, byte b of a row holds bits 8b+1 .. 8b+8, zero padded."""
    bits01 = np.asarray(bits01, np.uint8)
    N, D = bits01.shape
    packed = np.zeros((N, int(np.ceil(D / 8))), np.uint8)
    for b in range(1, D + 1):                                                       # :639-643
        byte_idx = int(np.ceil(b / 8))
        bit_pos = 8 - (b - 1) % 8
        packed[:, byte_idx - 1] |= bits01[:, b - 1] * np.uint8(2 ** (bit_pos - 1))
    return packed


def match_features(F1, F2, match_threshold=3.5, max_ratio=0.6, unique=True):
    """matchFeaturesScratch(F1, F2, 'Method', 'Exhaustive', ...) end to end (:81-126, :169-215).
    uint8 inputs are packed binary descriptors (binaryFeatures.Features)."""
    F1, F2 = np.asarray(F1), np.asarray(F2)
    if F1.dtype == np.uint8:
        if F1.size == 0 or F2.size == 0:                                            # :84-88
            return np.zeros((0, 2), np.uint32), np.zeros(0, np.float32)
        n_bits = F1.shape[1] * 8
        idx2, d1, d2 = nearest2_hamming(F1, F2)
        db, ds = hamming_percent(d1, d2, n_bits)
        return ratio_threshold_unique(idx2, db, ds, F1.shape[0], F2.shape[0], True, match_threshold, max_ratio, unique)
    A, B = F1.astype(np.float32), F2.astype(np.float32)
    if np.max(np.abs(A)) > 2 or np.max(np.abs(B)) > 2:                              # :105-110
        A, B = normalize_pairwise(A), normalize_pairwise(B)
    idx2, d1, d2 = nearest2_ssd(A, B)
    return ratio_threshold_unique(idx2, d1, d2, A.shape[0], B.shape[0], False, match_threshold, max_ratio, unique)


def select_partners(counts, m):
    """PP/imageMatching/imageMatching.m:75-100.  counts[i, j] = size(matchesAll{i+1,j+1}, 1).
    Returns (candidatePairs bool [n x n], IuptriIdx 1-based column-major linear indices)."""
    C = np.asarray(counts, np.int64)
    n = C.shape[0]
    sym = C + C.T                                                                   # :82
    sym[np.arange(n), np.arange(n)] = 0                                             # :83
    order = np.argsort(-sym, axis=1, kind="stable")                                 # :86 descending, stable
    top = order[:, :min(m, n - 1)] if n > 1 else order[:, :0]                       # :87
    cand = np.zeros((n, n), bool)
    for r in range(n):                                                              # :90-92
        cand[r, top[r]] = True
    cand = cand | cand.T                                                            # :95
    cand = np.triu(cand, 1)                                                         # :96
    lin = np.flatnonzero(cand.T.reshape(-1)) + 1                                    # :99 find(): column-major
    return cand, lin


def knn_bruteforce(train, query, k):
    """Output contract of flann_knn_win (PP/mex/flann_knn.cpp:193-253) with an exhaustive search: [Fq x k] uint32
    1-based indices in ascending distance (ties -> lower index), single distances (squared L2, or Hamming bit
    counts); missing neighbours 0 / +inf.  Float distances are evaluated in double and rounded to single, so they
    may differ from FLANN's functor in the last bit: use for inputs without near-ties."""
    train, query = np.asarray(train), np.asarray(query)
    if train.dtype == np.uint8:
        Dm = np.unpackbits(query[:, None, :] ^ train[None, :, :], axis=2).sum(axis=2).astype(np.float64)
    else:
        diff = query.astype(np.float64)[:, None, :] - train.astype(np.float64)[None, :, :]
        Dm = (diff * diff).sum(axis=2)
    order = np.argsort(Dm, axis=1, kind="stable")[:, :k]
    idx = np.zeros((query.shape[0], k), np.uint32)
    dist = np.full((query.shape[0], k), np.inf, np.float32)
    kk = order.shape[1]
    idx[:, :kk] = order + 1
    dist[:, :kk] = np.take_along_axis(Dm, order, axis=1).astype(np.float32)
    return idx, dist


def feature_matching_global(desc_list, k, ratio_thr):
    """featureMatchingGlobal.m:48-161 end to end with an exhaustive kNN.  Returns the CSR form of the cell used by
    the oracle and the library: pair_ptr [n*n + 1] over column-major cells, rows [M x 2] uint32."""
    n = len(desc_list)
    counts = [d.shape[0] for d in desc_list]
    nonempty = [d for d in desc_list if d.shape[0]]
    pair_ptr = np.zeros(n * n + 1, np.int64)
    if not nonempty:
        return pair_ptr, np.zeros((0, 2), np.uint32)
    pooled = np.concatenate(nonempty)
    if pooled.dtype != np.uint8:
        pooled = normalize_global(pooled)
    idx, dist = knn_bruteforce(pooled, pooled, k)
    cells = global_filter_and_scatter(idx, dist, counts, ratio_thr)
    rows = []
    for j in range(1, n + 1):
        for i in range(1, n + 1):
            c = (i - 1) + (j - 1) * n
            m = cells.get((i, j))
            pair_ptr[c + 1] = pair_ptr[c] + (0 if m is None else len(m))
            if m is not None:
                rows.append(m)
    rows = np.concatenate(rows).astype(np.uint32) if rows else np.zeros((0, 2), np.uint32)
    return pair_ptr, rows
