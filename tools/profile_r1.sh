#!/bin/bash
# Round-1 profiling pass (run under gpurun from the repo root): launch list of one short bench
# run + one --set full capture of the grouped GEMM kernel and the gather kernel.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 200 -c 6 \
    -o gpurun_out/gemm_r1 -f python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 20 -c 2 \
    -o gpurun_out/gather_r1 -f python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out
