#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fixture_parity" 2>&1 | grep "update =\|passed\|failed\|rror" | cut -c1-100
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
for a in tqc sac ddpg; do
  timeout 600 python bench.py --algo $a --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2i_bench_${a}.json
  python -c "
import json
d=json.load(open('gpurun_out/r2i_bench_${a}.json'))
print('$a value %.0f us/step %.1f e2e %.0f gemm %.1f simt %.1f launches %s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['roofline']['gemm_us_per_update'], d['roofline']['simt_us_per_update'], d.get('gpu_launches')))"
done
bash tools/gpu_ncu_list.sh tqc r2i 640 36
