"""C5-family pairwise sweep on N GPUs of one box (BASELINE.json config 5: "pairwise-block sweep at 2/4/8 GPUs").
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/time_pairwise_multi.py [n_images=300] [kp=4096] [iters=3]
Every rank computes every N-th image pair of the column-major pair list (aps_feature_matching_pairwise_shard, the
reference's parfor schedule, featureMatchingPairwise.m:48-59) and the per-pair lists are exchanged (all_gather_object):
strong scaling of a fixed set.  Time = max over ranks between barriers, pairs = sum_{i<j} N_i * N_j.
NOT YET RUN on a multi-GPU box (written after round 1's GPU budget was spent)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = ge.load_package()
from importlib import import_module  # noqa: E402

host = import_module(pkg.__name__ + ".host")
ctx = pkg.Context(local)
desc, c = pkg.synth.make_config(5, n=n, kp=kp)   # same seed on every rank
inp = {"Matchingmethod": "Exhaustive", "Matchingthreshold": 1.5, "Ratiothreshold": 0.7}
pairs = float(sum(desc[i].shape[0] * desc[j].shape[0] for j in range(n) for i in range(j)))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


best = None
for it in range(iters):
    barrier()
    t0 = time.perf_counter()
    m = pkg.multigpu.pairwise_matching_sharded(host, inp, desc, n, rank, world, dist if world > 1 else None, ctx=ctx)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    best = float(dt[0]) if best is None else min(best, float(dt[0]))
if rank == 0:
    rows = sum(m[i][j].shape[0] for j in range(n) for i in range(j))
    print(f"pairwise C5-family n={n} kp={kp} on {world} GPU(s): {pairs:.3e} descriptor pairs in {best * 1e3:.1f} ms "
          f"-> {pairs / best:.3e} pairs/s ({rows} match rows, merged on every rank)")
if world > 1:
    dist.destroy_process_group()
