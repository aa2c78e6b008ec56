"""Times featureMatchingPairwise (batched pipeline) on one GPU.  usage: time_pairwise.py cid n kp"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
cid, n, kp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
desc, c = pkg.synth.make_config(cid, n=n, kp=kp)
cells = [pkg.binaryFeatures(d) for d in desc] if c["kind"] == "orb" else desc
inp = {"Matchingmethod": "Exhaustive", "Matchingthreshold": 10.0 if c["kind"] == "orb" else 1.5, "Ratiothreshold": 0.7}
ctx = pkg._lib.default_context()
if os.environ.get("APS_PAIR_EPILOGUE"):   # 1 = branch-free segment selection (aps_ctx_set_pairwise_epilogue)
    ctx.set_pairwise_epilogue(int(os.environ["APS_PAIR_EPILOGUE"]))
pairs = sum(desc[i].shape[0] * desc[j].shape[0] for j in range(n) for i in range(j))
import ctypes as C

from importlib import import_module
host = import_module(pkg.__name__ + ".host")
for it in range(3):
    t0 = time.perf_counter()
    m = pkg.featureMatchingPairwise(inp, cells, n)
    dt = time.perf_counter() - t0
    if it == 2 and c["kind"] != "orb":  # the C-ABI call alone (host buffers in, CSR lists out)
        _, first, mats, counts, D, is_bin = host._describe(cells, n)
        ptrs, cnt, layout, keep = host._desc_args(mats, counts)
        h = C.c_void_p()
        t1 = time.perf_counter()
        host.check(host.lib().aps_feature_matching_pairwise_shard(ctx.handle, ptrs, cnt, n, int(D), host.APS_F32, layout,
                                                                  inp["Matchingthreshold"], inp["Ratiothreshold"], 0, 1, C.byref(h)))
        print(f"  C ABI call alone: {(time.perf_counter() - t1) * 1e3:.1f} ms")
        host.lib().aps_matchlist_free(h)
    rows = sum(m[i][j].shape[0] for j in range(n) for i in range(j))
    print(f"config {cid} n={n} kp={kp}: {n*(n-1)//2} image pairs, {pairs:.3e} descriptor pairs in {dt*1e3:.1f} ms "
          f"-> {pairs/dt:.3e} pairs/s end to end ({rows} match rows) {ctx.last_stats()}")
