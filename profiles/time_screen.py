"""Times the pairwise screen kernel alone (CUDA events inside the library): python profiles/time_screen.py [n] [kp] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
desc, c = pkg.synth.make_config(5, n=n, kp=kp)
D = desc[0].shape[1]
ctx = pkg.Context(0)
ctx.enable_timing(True)
plan = pkg.PairwisePlan(ctx, [d.shape[0] for d in desc], D, False)
plan.upload(desc)
plan.prepare()
pairs = float(sum(desc[i].shape[0] * desc[j].shape[0] for j in range(n) for i in range(j)))
best = None
for _ in range(reps):
    ctx.tc_time()
    plan.match(1.5, 0.7)
    ms, launches = ctx.tc_time()
    best = ms if best is None else min(best, ms)
print(f"variant={os.environ.get('APS_SCREEN_VARIANT', '0')}: tensor launches {launches}, {best:.2f} ms for {pairs:.3e} pairs "
      f"-> {2 * D * pairs / best / 1e9:.0f} TFLOP/s (screen + exact-stage tensor pass)")
