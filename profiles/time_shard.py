"""Times ONE rank's share of the weak-scaling bench on one GPU (no NCCL): the knn stage for query rows of rank
`rank` of `world` at the bench's image count for that world size.  usage: time_shard.py world [rank]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
world = int(sys.argv[1])
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = {1: 20, 2: 28, 4: 40, 8: 57}[world]
desc, c = pkg.synth.make_config(2, n=n, kp=8192)
torch.cuda.set_device(0)
ctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
plan = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
plan.upload(desc)
q0, q1 = pkg.multigpu.shard_bounds(plan.F, world)[rank]
plan.prepare()
for it in range(4):
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.knn(q0, q1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"world {world} rank {rank}: F={plan.F} rows [{q0},{q1}) knn {ms:.2f} ms -> {(q1-q0)*plan.F/ms*1e3:.3e} pairs/s per GPU")
