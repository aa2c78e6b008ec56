"""The MEX gateways in mex/ (drop-in names flann_knn_win, nearest2HammingExhaustive{,OMP}MEX and the batched
aps_featureMatching_mex / aps_imageMatching_mex / aps_matchFeatures_mex) compiled against the mex shim and driven by mex/shim_driver.cpp.
CPU: argument validation raises the reference's error identifiers and a missing GPU raises
apsmatch:nogpu (no CPU fallback).  GPU: the same binaries run a small real call."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GATES = ["gate_flann", "gate_hamming", "gate_hamming_omp", "gate_batched", "gate_imatch", "gate_matchf"]


def _build():
    if not os.path.exists(os.path.join(ROOT, "automaticpanoramicimagestitching-autopanostitch-matlab_b200", "libapsmatch.so")):
        pytest.skip("libapsmatch.so not built")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "mex"), "shimcheck"], check=True, capture_output=True)


@pytest.mark.parametrize("gate", GATES)
def test_gateway_argument_checks_cpu(gate):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    _build()
    r = subprocess.run([os.path.join(ROOT, "mex", "_build", gate)], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("gate", GATES)
def test_gateway_real_call_gpu(gate):
    _build()
    r = subprocess.run([os.path.join(ROOT, "mex", "_build", gate), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


# ---- real data through the gateways (file mode of mex/shim_driver.cpp), compared with the oracle ------------------
def _run_gate(gate, mats, scalars, tmp_path):
    """mats: list of (cls, array) with cls 0 single / 1 uint8 / 2 double / 4 uint8 inside a binaryFeatures object."""
    import numpy as np

    _build()
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.int32(len(mats)).tobytes())
        for cls, a in mats:
            a = np.asarray(a)
            f.write(np.array([cls, a.shape[0], a.shape[1]], np.int32).tobytes())
            f.write(np.asfortranarray(a).tobytes(order="F"))      # column-major, as mxGetData hands it over
        f.write(np.int32(len(scalars)).tobytes())
        f.write(np.asarray(scalars, np.float64).tobytes())
    r = subprocess.run([os.path.join(ROOT, "mex", "_build", gate), "file", fin, fout], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    raw = open(fout, "rb").read()
    nout = int(np.frombuffer(raw, np.int32, 1)[0])
    pos, outs = 4, []
    dt = {0: np.float32, 1: np.uint8, 2: np.float64, 3: np.uint32}
    for _ in range(nout):
        cls, rows, cols = (int(v) for v in np.frombuffer(raw, np.int32, 3, pos))
        pos += 12
        if cls < 0:
            outs.append(None)
            continue
        n = rows * cols * np.dtype(dt[cls]).itemsize
        outs.append(np.frombuffer(raw, dt[cls], rows * cols, pos).reshape((rows, cols), order="F"))
        pos += n
    return outs


def _cells_equal(outs, ref_cells, n):
    o = 0
    for j in range(n):
        for i in range(j):
            exp, got = ref_cells.get((i, j)), outs[o]
            o += 1
            if exp is None or len(exp) == 0:
                assert got is None or got.shape[0] == 0, (i, j)
            else:
                assert got is not None and got.shape == exp.shape and (got == exp).all(), (i, j)


@pytest.mark.gpu
def test_gateway_flann_knn_matches_the_oracle(tmp_path):
    import numpy as np

    from oracle import oracle

    rng = np.random.default_rng(5)
    T = oracle.normalize_rows_global(rng.integers(0, 256, (3000, 128)).astype(np.float32))
    T[100] = T[7]                                                 # duplicate row: self need not be first
    idx, dist = _run_gate("gate_flann", [(0, T)], [4], tmp_path)
    ri, rd = oracle.knn_l2(T, T, 4)
    assert idx.dtype == np.uint32 and dist.dtype == np.float32 and (idx == ri).all() and (dist == rd).all()
    Q = oracle.normalize_rows_global(rng.integers(0, 256, (257, 128)).astype(np.float32))
    idx, dist = _run_gate("gate_flann", [(0, T), (0, Q)], [3], tmp_path)
    ri, rd = oracle.knn_l2(T, Q, 3)
    assert (idx == ri).all() and (dist == rd).all()
    B = rng.integers(0, 256, (2000, 32), dtype=np.uint8)
    idx, dist = _run_gate("gate_flann", [(1, B)], [4, 1], tmp_path)   # 'bf'
    ri, rd = oracle.knn_hamming(B, B, 4)
    assert (idx == ri).all() and (dist == rd).all()


@pytest.mark.gpu
@pytest.mark.parametrize("gate", ["gate_hamming", "gate_hamming_omp"])
def test_gateway_hamming_matches_the_oracle(gate, tmp_path):
    import numpy as np

    from oracle import oracle

    rng = np.random.default_rng(6)
    A, B = rng.integers(0, 256, (1500, 32), dtype=np.uint8), rng.integers(0, 256, (1700, 32), dtype=np.uint8)
    B[5] = A[9]
    i2, d1, d2 = _run_gate(gate, [(1, A), (1, B)], [], tmp_path)
    r2, e1, e2 = oracle.nearest2_hamming(A, B)
    assert (i2.ravel() == r2).all() and (d1.ravel() == e1).all() and (d2.ravel() == e2).all()


@pytest.mark.gpu
def test_gateway_batched_global_and_pairwise_match_the_oracle(tmp_path):
    """float cells, binaryFeatures objects, and PLAIN uint8 matrices -- which the reference treats as float
    descriptors (featureMatchingGlobal.m:56,76-84: binary only by class), so the gateway must too."""
    import numpy as np

    from oracle import oracle

    pkg = __import__("__graft_entry__").load_package()
    desc, c = pkg.synth.make_config(1, n=4, kp=700)
    outs = _run_gate("gate_batched", [(0, d) for d in desc], [0, c["k"], c["ratio"], 0], tmp_path)
    _cells_equal(outs, oracle.feature_matching_global(desc, c["k"], c["ratio"])["cells"], len(desc))
    # doubles are cast to single like single(allDesc)
    outs = _run_gate("gate_batched", [(2, d.astype(np.float64)) for d in desc], [0, c["k"], c["ratio"], 0], tmp_path)
    _cells_equal(outs, oracle.feature_matching_global(desc, c["k"], c["ratio"])["cells"], len(desc))
    # plain uint8 matrices: float semantics (cast, L2-normalise, squared L2), NOT Hamming
    u8 = [d.astype(np.uint8) for d in desc]
    outs = _run_gate("gate_batched", [(1, d) for d in u8], [0, c["k"], c["ratio"], 0], tmp_path)
    _cells_equal(outs, oracle.feature_matching_global([d.astype(np.float32) for d in u8], c["k"], c["ratio"])["cells"], len(desc))
    # binaryFeatures objects: Hamming
    orb, cb = pkg.synth.make_config(4, n=3, kp=600)
    outs = _run_gate("gate_batched", [(4, d) for d in orb], [0, cb["k"], cb["ratio"], 1], tmp_path)
    _cells_equal(outs, oracle.feature_matching_global(orb, cb["k"], cb["ratio"])["cells"], len(orb))
    # pairwise
    kz, ck = pkg.synth.make_config(5, n=4, kp=500)
    outs = _run_gate("gate_batched", [(0, d) for d in kz], [1, 1.5, 0.6], tmp_path)
    ref = oracle.feature_matching_pairwise(kz, 1.5, 0.6)
    _cells_equal(outs, ref["cells"] if isinstance(ref, dict) else ref, len(kz))
    # 'Approximate': aps_method 1 'subsetpdist2' (the inputs.m default) and 3 'pca2nn' through the same gateway
    for method, fn in ((1, lambda a, b: oracle.match_features_method(a, b, 1.5, 0.6, "subsetpdist2")),
                       (3, lambda a, b: oracle.match_features_pca(a, b, 1.5, 0.6))):
        outs = _run_gate("gate_batched", [(0, d) for d in kz], [1, 1.5, 0.6, method], tmp_path)
        cells = {(i, j): fn(kz[i], kz[j])[0] for j in range(len(kz)) for i in range(j)}
        _cells_equal(outs, cells, len(kz))
