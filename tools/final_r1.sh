#!/bin/bash
# Final evidence pass of the round (one GPU): tests, every algorithm's bench line, the reference arm,
# and the --set full capture of the 12 GEMM launches of one DDPG update.
mkdir -p gpurun_out
(python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r1b_bench_ddpg.json 2> gpurun_out/bench.err
for a in td3 sac tqc; do python bench.py --algo $a --no-cpu-baseline > gpurun_out/r1b_bench_$a.json 2>> gpurun_out/bench.err; done
python bench.py --impl reference --steps 300 --warmup 10 > gpurun_out/r1b_bench_reference_arm.json 2>> gpurun_out/bench.err
B="python bench.py --steps 20 --warmup 10 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 120 -c 12 \
    -o gpurun_out/gemm_r1b -f $B > gpurun_out/ncu_gemm.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__cluster_dim_x,sm__cycles_active.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum"
ncu -i gpurun_out/gemm_r1b.ncu-rep --page raw --csv --metrics $M > gpurun_out/r1b_ncu_gemm_raw.csv 2> /dev/null
python - <<'PY'
import csv
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "sm__cycles_active.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open("gpurun_out/r1b_ncu_gemm_raw.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = [rows[hdr].index(c) for c in keep if c in rows[hdr]]
with open("gpurun_out/r1b_ncu_gemm_kernel_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows[hdr:]:
        if len(r) >= len(rows[hdr]):
            w.writerow([r[c] for c in cols])
PY
rm -f gpurun_out/r1b_ncu_gemm_raw.csv gpurun_out/gemm_r1b.ncu-rep
python tools/stage_profile.py > gpurun_out/r1b_stage_costs_ddpg.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r1b_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -3 gpurun_out/bench.err
