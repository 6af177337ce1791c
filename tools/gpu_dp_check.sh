#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_dp.py tests/test_gpu_distrib.py -x -q -s 2>&1 | grep "passed\|failed\|dp2\|rror\|Timeout" | tail -12) > gpurun_out/r2_pytest_gpu_2gpu.log 2>&1
cat gpurun_out/r2_pytest_gpu_2gpu.log
for a in ddpg tqc; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --algo $a --gpus 2 --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_${a}_n2.json
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_${a}_n2.json'))
print('$a N=2 value %.0f us/step %.1f e2e %.0f dp %s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d.get('dp_parity',{}).get('l2')))"
done
