#!/bin/bash
# Round-1 (second pass) profiling: launch list + full capture of the GEMM kernel after the v2 rewrite.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 210 --csv \
    --log-file gpurun_out/launches_r1b.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 210 -c 4 \
    -o gpurun_out/gemm_r1b -f python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
