#!/usr/bin/env python
"""bench.py -- descriptor-pairs/sec of the global feature-matching path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU arithmetic (oracle port)

A "step" is one pass of the hot path (K1 normalise/convert -> K2 tcgen05 candidates -> K3 exact
re-rank -> K5 ratio filter + compaction into match lists) over one synthetic descriptor set:
  N=1 : BASELINE.json configs[1] -- 20 images x 8192 SIFT-128 keypoints, global k=4, ratio 0.8
  N>1 : weak scaling -- round(20*sqrt(N)) images x 8192 keypoints, so that pairs per GPU (F^2/N) stay
        fixed; query rows are sharded across ranks, every rank holds all train descriptors,
        per-query match records are exchanged with NCCL and compacted.
`value`  : F^2 ordered descriptor pairs / device time of the step, inputs resident in HBM.
`e2e`    : same metric through the public call with pinned HOST buffers (H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "descriptor_pairs_per_sec"
UNIT = "pairs/s"
KP, D, KNN, RATIO = 8192, 128, 4, 0.8
IMAGES_FOR_GPUS = {1: 20, 2: 28, 4: 40, 8: 57}


def workload(n_gpus):
    n_img = IMAGES_FOR_GPUS.get(n_gpus, int(round(20 * n_gpus ** 0.5)))
    return n_img, f"C2-family: {n_img} images x {KP} SIFT-128 f32 keypoints, global exhaustive k={KNN}, ratio {RATIO}"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"burst": d.get("bf16_tflops"), "sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half of the samples = under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(desc, target_seconds, threads_note=True):
    """Times the oracle's exact float kNN (+ filter) on a bounded query sample against the full train set."""
    from oracle import oracle

    X = oracle.normalize_rows_global(np.concatenate(desc))
    F = X.shape[0]
    counts = np.array([d.shape[0] for d in desc], np.int64)
    cal = min(F, 256)
    t0 = time.perf_counter()
    oracle.knn_l2(X, X[:cal], KNN)
    rate = cal * F / (time.perf_counter() - t0)
    nq = int(min(F, max(cal, rate * target_seconds / F)))
    t0 = time.perf_counter()
    idx, dist = oracle.knn_l2(X, X[:nq], KNN)
    dt = time.perf_counter() - t0
    return {"value": nq * F / dt, "seconds": dt, "queries": nq, "F": F, "cores": oracle.num_threads()}


def flann_kdtree_rate(X, exact_idx1, q0, nq):
    """The reference's DEFAULT global float engine (PP/mex/flann_knn.cpp:229-234: cv::flann::Index, KDTree(4),
    knnSearch(checks=32)) through the OpenCV python module of this image -- approximate and randomised, so it is
    reported beside the exact arm, never as the parity oracle.  Returns None when cv2 is not importable."""
    try:
        import cv2
    except Exception:
        return None
    F = X.shape[0]
    t0 = time.perf_counter()
    index = cv2.flann_Index(X, dict(algorithm=1, trees=4))
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    idx, _ = index.knnSearch(X[q0:q0 + nq], KNN, params=dict(checks=32))
    t_search = time.perf_counter() - t0
    hits = sum(len(set(idx[r].tolist()) & set((exact_idx1[r] - 1).tolist())) for r in range(nq))
    total_s = t_build + t_search * F / nq
    return {"engine": f"cv2 {cv2.__version__} flann_Index KDTree(trees=4), knnSearch(checks=32); reference pins OpenCV 4.12",
            "value": F * float(F) / total_s, "unit": UNIT + " (F^2 / time: pair-equivalents of an approximate search)",
            "cores": 1, "build_s": t_build, "search_s_scaled": t_search * F / nq,
            "recall_at_k_vs_exact": hits / float(nq * KNN),
            "sample": f"index over all {F} rows, {nq} queries searched, search time scaled linearly"}


def blas_exhaustive_rate(X, nq):
    """The reference's own EXHAUSTIVE float arithmetic (PP/featureMatching/matchFeaturesScratch.m:343-358,
    nearest2SSDExhaustive: blocks of floor(1e7/N2) query rows, a2 + b2' - 2*A*B' by sgemm, min / mask / min) restated
    with numpy on this box's BLAS threads -- MATLAB would run the same three steps on MKL.  Extended here to the k = 4
    smallest per row (argpartition) so that it does the work of one global step.  Not bit-compatible with the oracle
    (sgemm summation order), so it is a timing arm only."""
    F = X.shape[0]
    block = max(1, int(1e7 // max(F, 1)))
    b2 = np.sum(X * X, axis=1, dtype=np.float32)
    t0 = time.perf_counter()
    for s0 in range(0, nq, block):
        A = X[s0:min(nq, s0 + block)]
        a2 = np.sum(A * A, axis=1, dtype=np.float32)
        D2 = (a2[:, None] + b2[None, :]) - np.float32(2) * (A @ X.T)
        part = np.argpartition(D2, KNN, axis=1)[:, :KNN]
        np.take_along_axis(D2, part, axis=1).sort(axis=1)
    dt = time.perf_counter() - t0
    return {"engine": "numpy sgemm + argpartition (nearest2SSDExhaustive's blocked GEMM form, k = %d)" % KNN,
            "value": nq * float(F) / dt, "unit": UNIT, "cores": os.cpu_count(),
            "sample": f"{nq} query rows x all {F} train rows in blocks of {block} rows ({dt:.1f} s), scaled linearly"}


def run_reference(args, rank, world):
    """--impl reference: the reference path's own CPU arithmetic on this box's host cores.
    MATLAB cannot run here and flann_knn.cpp needs OpenCV C++ (absent), so this is the oracle port
    (exact search with FLANN's own L2 functor, bit-identical distances to OpenCV's) with all OpenMP threads."""
    if rank != 0:
        return
    pkg = __import__("__graft_entry__").load_package()
    from oracle import oracle

    oracle.build(ref=False)
    n_img, wl = workload(args.gpus)
    desc, _ = pkg.synth.make_config(2, n=n_img, kp=KP)
    X = oracle.normalize_rows_global(np.concatenate(desc))
    F = X.shape[0]
    counts = np.array([d.shape[0] for d in desc], np.int64)
    cal = 256
    t0 = time.perf_counter()
    oracle.knn_l2(X, X[:cal], KNN)
    rate = cal * F / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)      # whole run within a few minutes
    nq = int(min(F, max(cal, rate * min(3.0, budget) / F)))
    times = []
    for it in range(args.warmup + args.steps):
        q0 = (it * nq) % max(1, F - nq)
        t0 = time.perf_counter()
        idx, dist = oracle.knn_l2(X, X[q0:q0 + nq], KNN)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = nq * F / (ms / 1e3)
    sample = f"{nq} query rows x all {F} train rows per step (exact kNN k={KNN}, FLANN L2 functor order), scaled linearly"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "F": F, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    approx = flann_kdtree_rate(X, idx, q0, nq)   # idx: the exact neighbours of the last timed sample
    if approx is not None:
        line["reference_default_engine"] = approx
    line["reference_exhaustive_blas"] = blas_exhaustive_rate(X, min(F, 4 * nq))
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = __import__("__graft_entry__").load_package()
    L = pkg._lib.lib()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(local_rank, stream=stream.cuda_stream)
    ctx.enable_timing(True)

    n_img, wl = workload(world)
    desc, _ = pkg.synth.make_config(2, n=n_img, kp=KP)   # same seed on every rank
    counts = [d.shape[0] for d in desc]
    F = sum(counts)
    pairs_total = float(F) * float(F)

    # pinned host copies (what a MEX gateway hands over), row-major
    import ctypes as C
    host_ptrs, host_views = [], []
    for d in desc:
        p = L.aps_host_alloc(d.nbytes)
        v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=d.shape)
        v[...] = d
        host_ptrs.append(p)
        host_views.append(v)

    plan = pkg.GlobalPlan(ctx, counts, D, False, KNN)
    plan.upload_pointers(host_ptrs)

    # query-row shard of this rank (contiguous blocks of 128-row tiles)
    mg = pkg.multigpu
    q0, q1 = mg.shard_bounds(F, world)[rank]
    rec = desc_dev = None
    if world > 1:
        rec = torch.as_tensor(mg.CudaView(plan.records_device(), (2 * F,), "<i4"), device="cuda")
        desc_dev = torch.as_tensor(mg.CudaView(plan.desc_device(), (F, D), "<f4"), device="cuda")

    def step_device():
        mg.global_matching_step(plan, RATIO, rank, world, dist, rec)

    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.tc_time()                                     # reset kernel-time accumulator
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.aps_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s0, s1 in ev:
        flush.zero_()
        barrier()
        s0.record(stream)
        step_device()
        s1.record(stream)
        barrier()
    launches = L.aps_launch_count() - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tc_ms, tc_launches = ctx.tc_time()
    stats = ctx.last_stats()
    if world > 1:
        t = torch.tensor([total_ms, tc_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, tc_ms_max = float(t[0]), float(t[1])
    else:
        tc_ms_max = tc_ms
    ms_per_step = total_ms / args.steps
    value = pairs_total / (ms_per_step / 1e3)

    # ---- end to end: pinned host buffers in, host match lists out -------------------------------
    def step_e2e():
        if world == 1:
            return pkg.featureMatchingGlobal({"k": KNN, "Ratiothreshold": RATIO}, host_views, n_img, ctx=ctx)
        # every rank uploads its own block of rows from pinned memory (all PCIe links in parallel), the blocks
        # are all-gathered over NVLink into the plan's pooled matrix
        mg.gather_descriptors(desc_dev, host_views, rank, world, dist, torch)
        step_device()
        return plan.download() if rank == 0 else None

    for _ in range(min(2, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        if world == 1:
            m_rows = sum(int(out[i][j].shape[0]) for j in range(n_img) for i in range(j) if out[i][j].ndim == 2)
        else:
            m_rows = int(out[3].shape[0])
        peaks = measured_peaks()
        # dominant kernel: k_knn_tc.  Algorithmic FLOPs per launch = 2*D * (rows of this rank) * F
        flops_per_launch = 2.0 * D * float(q1 - q0) * float(F)
        tc_avg_ms = tc_ms_max / max(1, tc_launches)
        achieved = flops_per_launch / (tc_avg_ms / 1e3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and world == 1:   # captured by ncu --set full on exactly this workload (profiles/)
            traffic = json.load(open(tp)).get("k_knn_tc_dram_bytes_per_launch")
        peak = peaks["sustained"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl, "F": F, "k": KNN, "ratio": RATIO, "pairs_per_step": pairs_total,
                       "arithmetic": "bf16 tcgen05 operands, f32 accumulate, exact f32 re-rank of the candidates",
                       "l2": "512 MB buffer written between timed steps (L2 flush)",
                       "sharding": f"query rows in {world} contiguous blocks; records exchanged by NCCL broadcast"
                                   + ("; e2e: every rank uploads its row block, NCCL all-gather of the blocks" if world > 1 else ""),
                       "engine": stats["engine"], "fallback_rows_last_step": stats["fallback_rows"],
                       "match_rows": m_rows},
            "roofline": {"bound": "tensor", "kernel": "k_knn_tc", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "peak_kind": f"bf16 sustained, {peaks['source']}",
                         "frac_of_burst": achieved / peaks["burst"], "kernel_ms": tc_avg_ms,
                         "kernel_share_of_step": tc_avg_ms / ms_per_step, "traffic": traffic},
            "e2e": {"value": pairs_total / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": F * D * 4, "d2h_bytes_per_step": (n_img * n_img + 1) * 8 + m_rows * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_rate(desc, 12.0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port",
                                    "sample": f"{cb['queries']} query rows x all {cb['F']} train rows "
                                              f"({cb['seconds']:.1f} s, exact kNN k={KNN} of the oracle, scaled linearly)"}
        print(json.dumps(line), flush=True)
    plan.close()
    for p in host_ptrs:
        L.aps_host_free(p)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)   # timing rule: at least 3 untimed warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
