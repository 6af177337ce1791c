#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5
for a in tqc; do
  OPRL_B200_DUMP_STAGES=1 timeout 300 python tools/stage_profile.py --algo $a > gpurun_out/r2c_stage_costs_$a.txt 2> gpurun_out/r2c_stage_plan_$a.txt
  grep -n "simt" -B0 gpurun_out/r2c_stage_plan_$a.txt | tail -8
  cat gpurun_out/r2c_stage_costs_$a.txt | tail -30
done
for a in tqc sac; do
  timeout 600 python bench.py --algo $a --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c_bench_$a.json
  python -c "
import json
d=json.load(open('gpurun_out/r2c_bench_$a.json'))
print('$a value %.0f us/step %.1f e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value']))"
done
for p in 0 1; do
OPRL_B200_ADAM_PATCH=$p timeout 600 python bench.py --algo ddpg --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ddpg patch=$p value %.0f us/step %.1f' % (d['value'], d['ms_per_step']*1e3))"
done
