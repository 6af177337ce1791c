#!/bin/bash
# Round-1 profiling pass (run under gpurun from the repo root):
#   1. launch list of one short bench run (cold-cache, serialised -> compare SHARES),
#   2. the same with caches left warm (--cache-control none) for per-kernel latency reading,
#   3. one --set full capture of the grouped GEMM kernel, the fused critic head and the gather kernel.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 10 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 190 --csv \
    --log-file gpurun_out/launches_r1.csv $B > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 380 -c 190 --csv \
    --log-file gpurun_out/launches_warm_r1.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 140 -c 14 \
    -o gpurun_out/gemm_r1 -f $B > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|critic_head_kernel|adam_kernel" -s 50 -c 5 \
    -o gpurun_out/simt_r1 -f $B > gpurun_out/ncu_simt.log 2>&1
ls -la gpurun_out
