#!/bin/bash
# One GPU round trip: parity tests, then the bench for each algorithm (fused dW0 on / off for DDPG).
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_ddpg.json 2> gpurun_out/bench_ddpg.err
OPRL_B200_DW0_GEMM=1 python bench.py --no-cpu-baseline > gpurun_out/bench_ddpg_nofuse.json 2>> gpurun_out/bench_ddpg.err
for a in "$@"; do
  python bench.py --algo $a --no-cpu-baseline > gpurun_out/bench_$a.json 2>> gpurun_out/bench_ddpg.err
done
tail -5 gpurun_out/bench_ddpg.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f (blocking %.0f) api %.0f us/upd %.1f gemm_us %.1f simt_us %.1f launches %s" % (
            d["value"], d["e2e"]["value"], d["e2e"].get("blocking_read_every_step", 0), d["api_loop"]["value"],
            d["ms_per_step"] * 1e3, d["roofline"]["gemm_us_per_update"], d["roofline"]["simt_us_per_update"], d["launches_per_update"]))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
