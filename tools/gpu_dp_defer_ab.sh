#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for d in 1 0; do
OPRL_B200_DW0_DEFER=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab_n${N}_defer$d.json
python - $N $d <<'PY'
import json, sys
n, d_ = sys.argv[1], sys.argv[2]
d = json.loads(open(f"gpurun_out/ab_n{n}_defer{d_}.json").read().strip().splitlines()[-1])
print("N=%s defer=%s value %.0f us/step %.1f e2e %.0f simt %.1f gemm %.1f" % (n, d_, d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d["roofline"]["simt_us_per_update"], d["roofline"]["gemm_us_per_update"]))
PY
done
