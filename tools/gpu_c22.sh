#!/bin/bash
for k in 148 112 96 64; do
for a in ddpg td3 sac tqc; do
  OPRL_B200_KSPLIT4_MAX_CTAS=$k timeout 600 python bench.py --algo $a --steps 1500 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('max4=$k $a value %.0f us/step %.2f gemm %.2f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['gemm_us_per_update']))"
done
done
