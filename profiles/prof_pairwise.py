"""Small driver used under ncu: the staged pairwise pipeline on one GPU (C5 family).
usage: python profiles/prof_pairwise.py [n_images=300] [kp=4096] [reps=1] [exhaustive|subsetpdist2|kdtree|pca2nn]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
method = sys.argv[4] if len(sys.argv) > 4 else "exhaustive"
desc, c = pkg.synth.make_config(5, n=n, kp=kp)
ctx = pkg.Context(0)
plan = pkg.PairwisePlan(ctx, [d.shape[0] for d in desc], desc[0].shape[1], False)
plan.set_method(method)
plan.upload(desc)
for _ in range(reps):
    ctx.synchronize()
    t0 = time.perf_counter()
    plan.prepare()
    pp, rows, met = plan.match(1.5, 0.7)
    ctx.synchronize()
    print(f"step {1e3 * (time.perf_counter() - t0):.1f} ms, {rows.shape[0]} match rows, {ctx.last_stats()}")
