#!/bin/bash
for k in 4 2 1; do
for a in ddpg td3; do
  OPRL_B200_KSPLIT=$k timeout 600 python bench.py --algo $a --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ksplit<=$k $a value %.0f us/step %.2f gemm %.2f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['gemm_us_per_update']))"
done
done
