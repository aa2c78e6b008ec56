"""Small driver used under ncu: runs the staged global pipeline a few times on one GPU.
usage: python profiles/prof_step.py [n_images] [kp] [reps] [config_id]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cid = int(sys.argv[4]) if len(sys.argv) > 4 else 2
desc, c = pkg.synth.make_config(cid, n=n, kp=kp)
ctx = pkg.Context(0)
is_bin = c["kind"] == "orb"
plan = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], desc[0].shape[1], is_bin, 4)
plan.upload(desc)
for _ in range(reps):
    plan.prepare()
    plan.knn()
    plan.filter(c["ratio"])
    plan.compact()
ctx.synchronize()
print("done", plan.F, ctx.last_stats())
