#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dp.py -x -q -s 2>&1 | grep "passed\|failed\|dp2\|rror" | tail -8) > gpurun_out/r2_pytest_gpu_dp_gate.log 2>&1
cat gpurun_out/r2_pytest_gpu_dp_gate.log
for g in 1 0; do
OPRL_B200_DP_GATE=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('gate $g: N=2 value %.0f us/step %.1f simt_us %.1f dp l2 %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['simt_us_per_update'], d['dp_parity']['l2']))
"
done
