"""Times the staged global pipeline on the BASELINE.json configs (one GPU).  usage: time_configs.py cid [n] [kp]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
cid = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
kp = int(sys.argv[3]) if len(sys.argv) > 3 else None
t0 = time.time()
desc, c = pkg.synth.make_config(cid, n=n, kp=kp)
print(f"config {cid}: {c['name']}  n={c['n']} kp={c['kp']} (generated in {time.time()-t0:.1f}s)")
ctx = pkg.Context(0)
ctx.enable_timing(True)
is_bin = c["kind"] == "orb"
plan = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], desc[0].shape[1], is_bin, 4)
plan.upload(desc)
F = plan.F
for it in range(3):
    ctx.synchronize()
    t = [time.perf_counter()]
    plan.prepare(); ctx.synchronize(); t.append(time.perf_counter())
    plan.knn(); ctx.synchronize(); t.append(time.perf_counter())
    plan.filter(c["ratio"]); plan.compact(); ctx.synchronize(); t.append(time.perf_counter())
    tc_ms, nl = ctx.tc_time()
    tot = t[-1] - t[0]
    D = desc[0].shape[1]
    print(f"  it{it}: prepare {1e3*(t[1]-t[0]):.2f} ms, knn {1e3*(t[2]-t[1]):.2f} ms (tc kernel {tc_ms:.2f} ms), filter+compact {1e3*(t[3]-t[2]):.2f} ms"
          f" | {F*F/tot:.3e} pairs/s" + (f" | tc {2*D*F*F/(tc_ms/1e3)/1e12:.0f} TFLOP/s" if tc_ms else f" | {F*F*D/ (t[2]-t[1]) /1e9:.0f} GB/s operand-equivalent"))
print("stats", ctx.last_stats())
m, _, pp, rows = plan.download()
print("match rows", len(rows))
