#!/bin/bash
# parity + properties + one bench line per algorithm (about two GPU-minutes)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_trainer.py -m gpu -x -q 2>&1 | tail -2
for a in ddpg td3 sac tqc; do
  timeout 600 python bench.py --algo $a --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$a value %.0f us/step %.2f gemm %.2f simt %.2f e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['gemm_us_per_update'], d['roofline']['simt_us_per_update'], d['e2e']['value']))"
done
