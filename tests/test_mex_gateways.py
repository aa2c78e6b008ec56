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
