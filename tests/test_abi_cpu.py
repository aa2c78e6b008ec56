"""CPU: the C-ABI library loads and exports every symbol include/apsmatch.h declares; without a
GPU the product fails loudly (no CPU fallback).  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "apsmatch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aps_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(aps):
    import ctypes

    if not os.path.exists(aps.library_path()):
        pytest.skip("libapsmatch.so not built (run __graft_entry__.build())")
    L = ctypes.CDLL(aps.library_path())
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/apsmatch.h but not exported"
    assert L.aps_abi_version() == 1


def test_binding_covers_the_header(aps):
    if not os.path.exists(aps.library_path()):
        pytest.skip("libapsmatch.so not built")
    L = aps._lib.lib()
    assert set(L._aps_symbols) == set(declared_symbols())


def test_no_cpu_fallback_without_gpu(aps):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(aps.library_path()):
        pytest.skip("libapsmatch.so not built")
    with pytest.raises(aps.ApsError) as e:
        aps.Context(0)
    assert e.value.code == 6 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "automaticpanoramicimagestitching-autopanostitch-matlab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle's", ""), f
