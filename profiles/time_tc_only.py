"""Runs ONLY the tensor candidate kernel (aps_debug_tc_scores without the score dump) on an n x n self-search:
python profiles/time_tc_only.py [rows] [D]   -- time it with  ncu --metrics gpu__time_duration.sum -k regex:k_knn_tc"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rng = np.random.default_rng(5)
X = rng.integers(0, 120, size=(n, D)).astype(np.float32)
ctx = pkg._lib.default_context()
L = pkg._lib.lib()
slots = 8 * L.aps_debug_tc_slots(ctx.handle, n, n)
cidx = np.zeros((n, slots), np.uint32)
csc = np.zeros((n, slots), np.float32)
for _ in range(3):
    pkg._lib.check(L.aps_debug_tc_scores(ctx.handle, X.ctypes.data, n, X.ctypes.data, n, D, 1, None, cidx.ctypes.data,
                                         csc.ctypes.data))
print("done", n, D, slots)
