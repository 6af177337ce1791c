#!/bin/bash
# scaling line at N GPUs (the driver's launch), kept under gpurun_out/r2_bench_ddpg_n$N.json
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_ddpg_n$N.json 2> gpurun_out/r2_bench_ddpg_n$N.err
tail -c 400 gpurun_out/r2_bench_ddpg_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2_bench_ddpg_n{n}.json").read().strip().splitlines()[-1])
print("N=%s value %.0f us/step %.1f e2e %.0f dp_parity l2 %s modes %s" % (n, d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d.get("dp_parity", {}).get("l2"), {k: (round(v["value"]) if isinstance(v, dict) else v) for k, v in d.get("scaling_modes", {}).items()}))
PY
