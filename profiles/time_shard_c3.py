"""One rank's share of the strong-scaling C3 job (100 x 10000 SIFT-128, F = 1e6) on ONE GPU, stage by stage with CUDA
events (no NCCL: the record exchange is the only stage missing).  usage: time_shard_c3.py [world] [rank]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
desc, c = pkg.synth.make_config(3)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = pkg.Context(0, stream=stream.cuda_stream)
plan = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
plan.upload(desc)
q0, q1 = pkg.multigpu.shard_bounds(plan.F, world)[rank]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
stages = [("prepare (K1)", lambda: plan.prepare()), ("knn (tensor + re-rank + fallback)", lambda: plan.knn(q0, q1)),
          ("filter", lambda: plan.filter(c["ratio"], q0, q1)), ("compact (all F records)", lambda: plan.compact())]
best = [1e9] * len(stages)
for it in range(5):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
    ctx.tc_time()
    ev[0].record(stream)
    for i, (_, fn) in enumerate(stages):
        fn()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    tc_ms, _ = ctx.tc_time()
    for i in range(len(stages)):
        best[i] = min(best[i], ev[i].elapsed_time(ev[i + 1]))
    print(f"it {it}: total {ev[0].elapsed_time(ev[-1]):.2f} ms, tensor kernel {tc_ms:.2f} ms, stats {ctx.last_stats()}")
for (name, _), b in zip(stages, best):
    print(f"{name:40s} {b:8.3f} ms")
