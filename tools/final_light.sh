#!/bin/bash
# short evidence refresh: test log, smoke, the bench lines of all four algorithms (no ncu pass)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/r2_pytest_gpu_1gpu.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_ddpg.json 2> gpurun_out/r2_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ddpg_steps20_warmup5.json 2>> gpurun_out/r2_bench.err
for a in td3 sac tqc; do
  python bench.py --algo $a --no-cpu-baseline > gpurun_out/r2_bench_$a.json 2>> gpurun_out/r2_bench.err
done
python - <<'PY'
import json
for a in ("ddpg", "ddpg_steps20_warmup5", "td3", "sac", "tqc"):
    d = json.loads(open(f"gpurun_out/r2_bench_{a}.json").read().strip().splitlines()[-1])
    print(a, "value %.0f us/step %.1f e2e %.0f frac %.5f cpu %s" % (d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value")))
PY
cat gpurun_out/r2_pytest_gpu_1gpu.log
