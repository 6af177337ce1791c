#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2_pytest_gpu.log 2>&1
cat gpurun_out/r2_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 500 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bench value %.0f us/step %.1f e2e %.0f api %.0f traffic %s src %s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['api_loop']['value'], d['roofline']['traffic'], d['roofline']['traffic_source']))
"
