#!/bin/bash
# Profiling pass (run under gpurun from the repo root; ONE GPU):
#   1. launch list of one short bench run (cold-cache, serialised -> compare SHARES, not absolutes),
#   2. the same with caches left warm (--cache-control none),
#   3. one --set full capture of the 12 grouped-GEMM launches of one DDPG update and of the SIMT kernels,
#   4. the in-situ stage costs (tools/stage_profile.py, not under ncu).
# Summaries are cut here (ncu is on the box) so only small CSVs travel back.
set -x
mkdir -p gpurun_out
TAG=${1:-r1b}
B="python bench.py --steps 20 --warmup 10 --no-cpu-baseline"
# one learner step = 1 gather + 15 update launches; skip the warm-up steps, keep 10 steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 160 --csv \
    --log-file gpurun_out/${TAG}_launches_ddpg_b256.csv $B > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 320 -c 160 --csv \
    --log-file gpurun_out/${TAG}_launches_ddpg_b256_warm.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 120 -c 12 \
    -o gpurun_out/gemm_${TAG} -f $B > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|critic_head_kernel|adam_kernel" -s 40 -c 4 \
    -o gpurun_out/simt_${TAG} -f $B > gpurun_out/ncu_simt.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__cluster_dim_x,sm__cycles_active.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum"
for k in gemm simt; do
  ncu -i gpurun_out/${k}_${TAG}.ncu-rep --page raw --csv --metrics $M > gpurun_out/${TAG}_ncu_${k}_raw.csv 2> gpurun_out/ncu_export_${k}.log
done
python - "$TAG" <<'PY'
import csv, sys
tag = sys.argv[1]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "sm__cycles_active.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for k, out in (("gemm", "gemm_kernel"), ("simt", "simt_kernels")):
    try:
        rows = list(csv.reader(open(f"gpurun_out/{tag}_ncu_{k}_raw.csv")))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        cols = [rows[hdr].index(c) for c in keep if c in rows[hdr]]
        with open(f"gpurun_out/{tag}_ncu_{out}_summary.csv", "w", newline="") as f:
            w = csv.writer(f)
            for r in rows[hdr:]:
                if len(r) >= len(rows[hdr]):
                    w.writerow([r[c] for c in cols])
    except Exception as ex:
        print("summary failed for", k, ex)
PY
python tools/stage_profile.py > gpurun_out/${TAG}_stage_costs_ddpg.txt 2>&1
rm -f gpurun_out/${TAG}_ncu_*_raw.csv
ls -la gpurun_out
