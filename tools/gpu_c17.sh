#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_properties.py tests/test_gpu_trainer.py tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -3
for a in ddpg td3 sac; do
timeout 600 python bench.py --algo $a --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$a value %.0f us/step %.2f e2e %.0f (%.2f us) blocking %.0f api %.0f last_loss %s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], 1e6/d['e2e']['value'], d['e2e']['blocking_read_every_step'], d['api_loop']['value'], d['e2e']['last_critic_loss']))"
done
