#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./build/gemm_selftest 2>&1 | grep -i "fails" | head -3
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fixture_parity" 2>&1 | grep "update =\|passed\|failed\|rror" | cut -c1-100
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -3
cp oprl_b200/liboprl_b200.so /tmp/new.so
for v in old new; do
  if [ $v = old ]; then cp build/liboprl_old.so oprl_b200/liboprl_b200.so; else cp /tmp/new.so oprl_b200/liboprl_b200.so; fi
  for a in ddpg td3; do
  timeout 600 python bench.py --algo $a --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v $a value %.0f us/step %.2f simt %.2f gemm %.2f launches %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['simt_us_per_update'], d['roofline']['gemm_us_per_update'], d.get('gpu_launches')))"
  done
done
cp /tmp/new.so oprl_b200/liboprl_b200.so
for a in tqc sac; do
  timeout 600 python bench.py --algo $a --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2h_bench_${a}.json
  python -c "
import json
d=json.load(open('gpurun_out/r2h_bench_${a}.json'))
print('$a value %.0f us/step %.1f e2e %.0f gemm %.1f simt %.1f launches %s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['roofline']['gemm_us_per_update'], d['roofline']['simt_us_per_update'], d.get('gpu_launches')))"
done
