#!/bin/bash
cp oprl_b200/liboprl_b200.so /tmp/new.so
for rep in 1 2; do
for v in old new; do
  if [ $v = old ]; then cp build/liboprl_old.so oprl_b200/liboprl_b200.so; else cp /tmp/new.so oprl_b200/liboprl_b200.so; fi
  for a in ddpg td3; do
  timeout 600 python bench.py --algo $a --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v $a value %.0f us/step %.2f simt %.2f gemm %.2f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['simt_us_per_update'], d['roofline']['gemm_us_per_update']))"
  done
done
done
cp /tmp/new.so oprl_b200/liboprl_b200.so
