#!/bin/bash
# One gpurun session: GPU tests, then the bench lines of every single-GPU config.  Outputs under gpurun_out/.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_session.sh tests bench'
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests)   timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/tests.log;;
    bench)   for c in c2 c3 c4 c5; do
               timeout 400 python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"; head -c 1800 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err; done;;
    record)  for c in c2 c3 c4 c5 c6; do
               timeout 400 python bench.py --config $c > gpurun_out/r2_bench_${c}_1gpu.json 2> gpurun_out/r2_bench_${c}_1gpu.err; echo "bench $c rc=$?"; head -c 700 gpurun_out/r2_bench_${c}_1gpu.json; echo; tail -2 gpurun_out/r2_bench_${c}_1gpu.err; done;;
    bench2)  timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; head -c 2500 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err;;
    ref)     timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; head -c 1500 gpurun_out/bench_ref.json;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/launches_bench.log 2>&1; echo "launches rc=$?";;
    launches_c5) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c5.csv python profiles/prof_pairwise.py 300 4096 1 > gpurun_out/launches_c5.log 2>&1; echo "launches_c5 rc=$?"; tail -2 gpurun_out/launches_c5.log;;
    ncu_screen) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair_screen$ -c 1 \
               --metrics sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,sm__inst_executed_pipe_tmem.sum \
               -o gpurun_out/prof_screen -f python profiles/prof_pairwise.py 120 4096 1 > gpurun_out/ncu_screen.log 2>&1; echo "ncu_screen rc=$?"; tail -3 gpurun_out/ncu_screen.log;;
    pairtests) timeout 900 python -m pytest tests -m gpu -x -q -k "pair or match_features or screen or c5 or gateway_batched" > gpurun_out/pairtests.log 2>&1; echo "pairtests rc=$?"; tail -5 gpurun_out/pairtests.log;;
    bench_c5) timeout 400 python bench.py --config c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; head -c 2200 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err;;
    ncu_all) # one full capture per kernel family of round 2 (profiles/r2_ncu_*.txt are made from these)
             timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_knn_tc.*int.6, .int.1" -c 1 --metrics sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,sm__inst_executed_pipe_tmem.sum -o gpurun_out/prof_tc6 -f python profiles/prof_step.py 20 8192 2 2 > gpurun_out/ncu_tc6.log 2>&1; echo "tc6 rc=$?"
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair_screen$ -c 1 --metrics sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,sm__inst_executed_pipe_tmem.sum -o gpurun_out/prof_screen2 -f python profiles/prof_pairwise.py 120 4096 1 > gpurun_out/ncu_screen2.log 2>&1; echo "screen rc=$?"
             timeout 600 ncu --set full --clock-control none -k regex:'k_prepare_norm|k_prepare_operands|k_rerank|k_global_filter|k_rank_per_image|k_scatter_rows|k_gather_train' -s 7 -c 7 -o gpurun_out/prof_aux2 -f python profiles/prof_step.py 20 8192 3 2 > gpurun_out/ncu_aux2.log 2>&1; echo "aux rc=$?"
             timeout 600 ncu --set full --clock-control none -k regex:k_knn_hamming -c 1 -o gpurun_out/prof_ham -f python profiles/prof_step.py 20 8192 1 4 > gpurun_out/ncu_ham.log 2>&1; echo "ham rc=$?";;
    ncu_tc6) timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_knn_tc.*int.6, .int.1" -c 1 --metrics sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,sm__inst_executed_pipe_tmem.sum -o gpurun_out/prof_tc6 -f python profiles/prof_step.py 20 8192 2 2 > gpurun_out/ncu_tc6.log 2>&1; echo "tc6 rc=$?";;
    ncu_tc)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_tc -s 2 -c 1 \
               --metrics sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,sm__inst_executed_pipe_tmem.sum \
               -o gpurun_out/prof_tc -f python profiles/prof_step.py 20 8192 2 2 > gpurun_out/ncu_tc.log 2>&1; echo "ncu_tc rc=$?"; tail -3 gpurun_out/ncu_tc.log;;
    ncu_aux) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_prepare_norm|k_prepare_operands|k_rerank|k_global_filter|k_rank_per_image|k_scatter_rows|k_gather_train' -s 8 -c 8 \
               -o gpurun_out/prof_aux -f python profiles/prof_step.py 20 8192 3 2 > gpurun_out/ncu_aux.log 2>&1; echo "ncu_aux rc=$?"; tail -3 gpurun_out/ncu_aux.log;;
    *) echo "unknown step $what";;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
