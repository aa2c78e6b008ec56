"""Where the end-to-end milliseconds of the one-call global path go (C2): python profiles/time_e2e.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
from importlib import import_module

host = import_module(pkg.__name__ + ".host")
L = pkg._lib.lib()
ctx = pkg.Context(0)
desc, c = pkg.synth.make_config(2)
views = []
for d in desc:
    p = L.aps_host_alloc(d.nbytes)
    v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=d.shape)
    v[...] = d
    views.append(v)
inp = {"k": 4, "Ratiothreshold": c["ratio"]}
for _ in range(3):
    pkg.featureMatchingGlobal(inp, views, len(views), ctx=ctx)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0)
N = 20
for _ in range(N):
    t0 = time.perf_counter(); n, first, mats, counts, D, is_binary = host._describe(views, len(views)); tick("describe", t0)
    t0 = time.perf_counter(); ptrs, cnt, layout, keep = host._desc_args(mats, counts); tick("desc_args", t0)
    h = C.c_void_p()
    t0 = time.perf_counter()
    L.aps_feature_matching_global(ctx.handle, ptrs, cnt, n, int(D), 0, layout, 4, c["ratio"], 0, C.byref(h)); tick("C call (H2D + kernels + D2H)", t0)
    t0 = time.perf_counter(); m = host._cells_from_matchlist(h, n); tick("cells", t0)
    t0 = time.perf_counter(); L.aps_matchlist_free(h); tick("free", t0)
# staged: separate the pieces of the C call
plan = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
ptr_list = [v.ctypes.data for v in views]
for _ in range(N):
    ctx.synchronize()
    t0 = time.perf_counter(); plan.upload_pointers(ptr_list); ctx.synchronize(); tick("  staged: upload", t0)
    t0 = time.perf_counter(); plan.prepare(); plan.knn(); plan.filter(c["ratio"]); plan.compact(); ctx.synchronize(); tick("  staged: device work", t0)
    t0 = time.perf_counter(); out = plan.download(); tick("  staged: download + cells", t0)
t0 = time.perf_counter()
for _ in range(N):
    p2 = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4); p2.close()
tick("  plan create + destroy", t0)
for k, v in T.items():
    print(f"{k:34s} {1e3 * v / N:7.3f} ms")
