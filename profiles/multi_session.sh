#!/bin/bash
# Multi-GPU session (run under gpurun --gpus N): bench arms of this repo on N GPUs.  Usage: bash profiles/multi_session.sh N [configs...]
N=$1; shift
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
for c in "$@"; do
  case $c in
    ref) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "ref rc=$?"; head -c 900 gpurun_out/bench_ref_${N}gpu.json;;
    *)   timeout 900 bash -c "$(declare -f run); N=$N; run --config $c" > gpurun_out/bench_${c}_${N}gpu.json 2> gpurun_out/bench_${c}_${N}gpu.err; echo "bench $c x$N rc=$?"; grep '^{' gpurun_out/bench_${c}_${N}gpu.json | head -c 2600; echo; tail -4 gpurun_out/bench_${c}_${N}gpu.err;;
  esac
done
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
