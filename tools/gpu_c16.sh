#!/bin/bash
cp oprl_b200/liboprl_b200.so /tmp/new.so
for rep in 1 2; do
for v in new nofence; do
  if [ $v = new ]; then cp /tmp/new.so oprl_b200/liboprl_b200.so; else cp build/liboprl_$v.so oprl_b200/liboprl_b200.so; fi
  timeout 600 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v value %.0f us/step %.2f e2e %.0f (%.2f us) blocking %.0f api %.0f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], 1e6/d['e2e']['value'], d['e2e']['blocking_read_every_step'], d['api_loop']['value']))"
done
done
cp /tmp/new.so oprl_b200/liboprl_b200.so
