"""CPU: live pinning of the oracle against oracle/_ref (the reference's MEX sources compiled
verbatim).  Skipped where _ref has not been built (it is built by __graft_entry__.build() in the
container that has /root/reference and travels to the GPU box as a prebuilt .so)."""
import numpy as np
import pytest


@pytest.mark.parametrize("omp", [False, True])
@pytest.mark.parametrize("shape", [(33, 47, 32), (128, 200, 64), (40, 40, 1), (7, 1, 32), (9, 0, 32), (0, 5, 32)])
def test_hamming_restatement_equals_reference_build(orc, omp, shape):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    n1, n2, nb = shape
    rng = np.random.default_rng(n1 * 1000 + n2)
    hi = 256 if nb > 1 else 8
    A = rng.integers(0, hi, (n1, nb), dtype=np.uint8)
    B = rng.integers(0, hi, (n2, nb), dtype=np.uint8)
    ref = orc.ref_nearest2_hamming(A, B, omp=omp)
    got = orc.nearest2_hamming(A, B)
    for r, g in zip(ref, got):
        assert np.array_equal(r, g, equal_nan=True)
