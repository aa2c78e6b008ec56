"""CPU: the oracle restatement against the committed golden vectors (tests/golden/make_golden.py
says where each vector comes from: the reference's own MEX sources compiled verbatim, and the
OpenCV engines flann_knn.cpp calls)."""
import numpy as np
import pytest

HAM = ["rand256", "ties", "n2is1", "n2is0", "dups512"]


@pytest.mark.parametrize("name", HAM)
def test_nearest2_hamming_matches_reference_mex(golden, orc, name):
    A, B = golden[f"ham2nn_{name}_A"], golden[f"ham2nn_{name}_B"]
    idx2, d1, d2 = orc.nearest2_hamming(A, B)
    assert np.array_equal(idx2, golden[f"ham2nn_{name}_idx2"])
    assert np.array_equal(d1, golden[f"ham2nn_{name}_d1"], equal_nan=True)
    assert np.array_equal(d2, golden[f"ham2nn_{name}_d2"], equal_nan=True)


@pytest.mark.parametrize("name", ["rand256", "ties", "self", "kgtF"])
def test_knn_hamming_matches_bfmatcher(golden, orc, name):
    idx, dist = orc.knn_hamming(golden[f"bfknn_{name}_T"], golden[f"bfknn_{name}_Q"], 4)
    assert np.array_equal(idx, golden[f"bfknn_{name}_idx"])
    assert np.array_equal(dist, golden[f"bfknn_{name}_dist"])


@pytest.mark.parametrize("name,k", [("sift_self", 4), ("kaze", 4), ("tail37", 3)])
def test_knn_l2_matches_flann_bits(golden, orc, name, k):
    """Indices AND float32 distance bits equal FLANN's L2 functor (squared distances)."""
    idx, dist = orc.knn_l2(golden[f"flann_{name}_T"], golden[f"flann_{name}_Q"], k)
    assert np.array_equal(idx, golden[f"flann_{name}_idx"])
    assert np.array_equal(dist.view(np.uint32), golden[f"flann_{name}_dist"].view(np.uint32))


def test_duplicates_self_not_first(golden):
    """with exact duplicates the self index need not be the first neighbour."""
    idx = golden["flann_sift_self_idx"]
    rows = np.arange(1, idx.shape[0] + 1)
    assert (idx[:, 0] != rows).any()
    assert ((idx == rows[:, None]).sum(1) == 1).all()
