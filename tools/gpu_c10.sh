#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fixture_parity" 2>&1 | grep "update =\|passed\|failed\|rror" | cut -c1-100
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -3
for a in tqc sac td3 ddpg; do
  OPRL_B200_DUMP_STAGES=1 timeout 300 python tools/stage_profile.py --algo $a > gpurun_out/r2e_stage_costs_$a.txt 2> gpurun_out/r2e_stage_plan_$a.txt
  tail -2 gpurun_out/r2e_stage_costs_$a.txt | head -1
done
for a in tqc sac; do
  for w in 0 1; do
  OPRL_B200_GEMM_PARTITION=$w timeout 600 python bench.py --algo $a --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2e_bench_${a}_p$w.json
  python -c "
import json
d=json.load(open('gpurun_out/r2e_bench_${a}_p$w.json'))
print('$a partition=$w value %.0f us/step %.1f e2e %.0f gemm %.1f launches %s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['roofline']['gemm_us_per_update'], d.get('gpu_launches')))"
  done
done
for a in ddpg td3; do
  timeout 600 python bench.py --algo $a --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$a value %.0f us/step %.2f simt %.2f gemm %.2f launches %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['simt_us_per_update'], d['roofline']['gemm_us_per_update'], d.get('gpu_launches')))"
done
