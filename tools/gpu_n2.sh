#!/bin/bash
# 2-GPU pass: the data-parallel parity tests, the 2-learner distributed run, then the scaling bench line with dp_parity
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_distrib.py -x -q -s 2>&1 | grep -v "^$" | tail -25) > gpurun_out/r2_pytest_gpu_2gpu.log 2>&1
cat gpurun_out/r2_pytest_gpu_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_ddpg_n2.json 2> gpurun_out/r2_bench_ddpg_n2.err
tail -c 600 gpurun_out/r2_bench_ddpg_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_ddpg_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.0f us/step %.1f e2e %.0f dp_parity %s scaling_modes %s" % (d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d.get("dp_parity"), d.get("scaling_modes")))
PY
