"""Soundness of the completeness rule behind the segment epilogue, checked on a numpy MODEL of the kernel and the re-rank
(CPU only; the CUDA code itself is covered by tests/test_gpu_parity.py::test_pairwise_segment_epilogue_variant).

Model (csrc/aps_knn_tc.cu, KCT == 3 / CS == 2, and csrc/aps_rerank.cu, tile mode):
  * the columns of a train image are cut into segments of SEG columns on the global grid; list `ch` of a row sees the
    segments with (segment index % 2) == ch;
  * per segment the two largest keys are offered; a list keeps the LEN largest offers, sorted, ties resolved in favour
    of the earlier offer;
  * rule: a column outside list s is bounded by the list's last entry, except the columns of a segment two of whose
    offers sit BEFORE the last entry: that segment is scanned exactly (at most one per list); a list that is not full
    bounds its outsiders by its last valid entry.
The property: every column that is in no list and in no scanned segment has a key <= the rule's bound W (in score
terms; the re-rank states it for distances, a decreasing function of the score)."""
import numpy as np
import pytest

SEG = 64


def model_lists(keys, t0, LEN):
    """keys: scores of one query row against the columns [t0, t0 + n) of the pool (global column = t0 + local).
    Returns per list ch: (values, global columns) of its LEN best offers, best first (NaN-free input)."""
    n = keys.shape[0]
    gcol = t0 + np.arange(n)
    seg = gcol // SEG
    out = []
    for ch in (0, 1):
        offers_v, offers_c = [], []
        for sgm in np.unique(seg):
            if sgm % 2 != ch:
                continue
            m = np.flatnonzero(seg == sgm)
            order = m[np.argsort(-keys[m], kind="stable")][:2]      # two best of the segment
            offers_v += list(keys[order])
            offers_c += list(gcol[order])
        offers_v, offers_c = np.array(offers_v), np.array(offers_c, np.int64)
        o = np.argsort(-offers_v, kind="stable")[:LEN] if offers_v.size else np.zeros(0, int)
        out.append((offers_v[o], offers_c[o]))
    return out


def rule_bound(lists, LEN):
    """(W as a SCORE bound, set of scanned segments).  Mirrors the W computation of k_rerank in tile mode."""
    W = -np.inf
    scanned = set()
    for vals, cols in lists:
        nv = len(vals)
        if nv == 0:
            continue                       # no selectable column: nothing outside this list
        if nv < LEN:
            w = vals[nv - 1]
        else:
            w = vals[LEN - 1]
            for i in range(LEN - 1):
                for j in range(i + 1, LEN - 1):
                    if cols[i] // SEG == cols[j] // SEG:
                        scanned.add(int(cols[i] // SEG))
        W = max(W, w)                       # distance W = min over lists  <=>  score bound = max over lists
    return W, scanned


@pytest.mark.parametrize("LEN", [3, 4])
def test_rule_bounds_every_outsider(LEN):
    rng = np.random.default_rng(100 + LEN)
    worst_slack = np.inf
    for trial in range(400):
        n = int(rng.choice([1, 2, 3, 5, 63, 64, 65, 129, 300, 1000]))
        t0 = int(rng.integers(0, 500))
        kind = trial % 4
        if kind == 0:
            keys = rng.standard_normal(n)
        elif kind == 1:
            keys = rng.integers(0, 4, n).astype(np.float64)                  # massive ties
        elif kind == 2:
            keys = rng.standard_normal(n)
            a = int(rng.integers(0, n))
            keys[a:a + 4] = 10.0                                              # a run of equal best columns in one place
        else:
            keys = -np.abs(rng.standard_normal(n))
            keys[rng.integers(0, n, min(n, 3))] = 5.0 + rng.random(min(n, 3))  # a few clear winners
        lists = model_lists(keys, t0, LEN)
        W, scanned = rule_bound(lists, LEN)
        inlist = set()
        for _, cols in lists:
            inlist |= set(int(c) for c in cols)
        for local in range(n):
            g = t0 + local
            if g in inlist or g // SEG in scanned:
                continue
            assert keys[local] <= W, (LEN, trial, n, t0, local, keys[local], W)
            worst_slack = min(worst_slack, W - keys[local])
    assert worst_slack >= 0


def test_three_entries_without_scan_would_be_unsound():
    """the case the scan exists for: the two best columns of the image sit in one segment together with the third best"""
    keys = np.full(256, -1.0)
    keys[[3, 4, 5]] = [9.0, 8.0, 7.0]           # one segment holds the three best columns
    lists = model_lists(keys, 0, 3)
    W, scanned = rule_bound(lists, 3)
    assert 0 in scanned                          # segment 0 must be scanned exactly ...
    assert keys[5] > W                           # ... because its third best beats the list's last entry
