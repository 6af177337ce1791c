#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --algo tqc --steps 6 --warmup 4 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 600 -c 70 --csv \
    --log-file gpurun_out/r2b_launches_tqc_warm.csv $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 70 --csv \
    --log-file gpurun_out/r2b_launches_tqc.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|gemm_kernel|tqc_loss" -s 560 -c 30 \
    -o gpurun_out/tqc_r2b -f $B > gpurun_out/ncu_tqc.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__cycles_active.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed"
ncu -i gpurun_out/tqc_r2b.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2b_ncu_tqc_raw.csv 2> gpurun_out/ncu_export.log
rm -f gpurun_out/tqc_r2b.ncu-rep
ls -la gpurun_out | tail -5
