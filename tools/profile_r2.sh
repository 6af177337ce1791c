#!/bin/bash
# Round-2 profiling pass (run under gpurun from the repo root; ONE GPU):
#   1. launch list of a short bench run (cold-cache, serialised -> compare SHARES, not absolutes) + warm-cache variant,
#   2. one --set full capture of the chain / GEMM / SIMT kernels of one DDPG update, raw-page CSV kept
#      (bench.py reads roofline.traffic from it),
#   3. the chain kernel's own clock64 profile (tools/chain_prof.py, not under ncu).
set -x
mkdir -p gpurun_out
TAG=${1:-r2}
B="python bench.py --steps 20 --warmup 10 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 160 --csv \
    --log-file gpurun_out/${TAG}_launches_ddpg_b256.csv $B > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 200 -c 160 --csv \
    --log-file gpurun_out/${TAG}_launches_ddpg_b256_warm.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|gemm_kernel|adam_kernel|gather_kernel" -s 80 -c 8 \
    -o gpurun_out/update_${TAG} -f $B > gpurun_out/ncu_update.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__cluster_dim_x,sm__cycles_active.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed.sum"
ncu -i gpurun_out/update_${TAG}.ncu-rep --page raw --csv --metrics $M > gpurun_out/${TAG}_ncu_update_raw.csv 2> gpurun_out/ncu_export.log
python tools/chain_prof.py ddpg > gpurun_out/${TAG}_chain_prof_ddpg.txt 2>&1
python tools/chain_prof.py td3 > gpurun_out/${TAG}_chain_prof_td3.txt 2>&1
ls -la gpurun_out | tail -20
