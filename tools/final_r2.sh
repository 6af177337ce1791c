#!/bin/bash
# Round-2 evidence pass on ONE GPU: test log, bench lines of all four algorithms (+ the driver-like short window, the
# reference arm and the two opt-in chain plans), then the ncu pass (tools/profile_r2.sh).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2_pytest_gpu_1gpu.log 2>&1
python bench.py > gpurun_out/r2_bench_ddpg.json 2> gpurun_out/r2_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ddpg_steps20_warmup5.json 2>> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
for a in td3 sac tqc; do
  python bench.py --algo $a --no-cpu-baseline > gpurun_out/r2_bench_$a.json 2>> gpurun_out/r2_bench.err
done
OPRL_B200_CHAIN=1 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_ddpg_plan_full_chain.json 2>> gpurun_out/r2_bench.err
OPRL_B200_CHAIN_ACTOR=1 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_ddpg_plan_actor_chain.json 2>> gpurun_out/r2_bench.err
bash tools/profile_r2.sh r2 > gpurun_out/r2_profile.log 2>&1
# the wide configurations: launch plan + in-situ stage costs, warm launch list, one --set full pass over a TQC update
for a in ddpg td3 sac tqc; do
  OPRL_B200_DUMP_STAGES=1 timeout 300 python tools/stage_profile.py --algo $a > gpurun_out/r2_stage_costs_$a.txt 2> gpurun_out/r2_stage_plan_$a.txt
done
for a in sac tqc; do bash tools/gpu_ncu_list.sh $a r2 500 70 > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|gemm_kernel|tqc_loss" -s 560 -c 30 \
    -o gpurun_out/tqc_r2 -f python bench.py --algo tqc --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_tqc.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,sm__cycles_active.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum"
ncu -i gpurun_out/tqc_r2.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2_ncu_tqc_update_raw.csv 2>> gpurun_out/ncu_export.log
rm -f gpurun_out/tqc_r2.ncu-rep
python tools/sass_summary.py > gpurun_out/r2_sass_kernels.txt 2>&1
tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "value %.0f us/step %.1f e2e %.0f launches %s frac %s cpu %s" % (
            d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d.get("launches_per_update"), r.get("frac"),
            (d.get("cpu_baseline") or {}).get("value")))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
cat gpurun_out/r2_pytest_gpu_1gpu.log
ls -la gpurun_out | grep r2_
