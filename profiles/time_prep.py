"""Times K1's norm pass alone (CUDA events): python profiles/time_prep.py [n_images] [kp]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
desc, c = pkg.synth.make_config(2, n=n, kp=kp)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = pkg.Context(0, stream=stream.cuda_stream)
plan = pkg.GlobalPlan(ctx, [d.shape[0] for d in desc], 128, False, 4)
plan.upload(desc)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    plan.prepare()
    e1.record(stream)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
F = plan.F
print(f"K1 (norm + operands + scale sort + tile bounds) F={F}: {best * 1e3:.1f} us; norm pass moves {F * 128 * 8 / 1e6:.0f} MB")
