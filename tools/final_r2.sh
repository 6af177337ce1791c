#!/bin/bash
# Round-2 evidence pass on ONE GPU: test log, bench lines of all four algorithms (+ the driver-like short window, the
# reference arm and the two opt-in chain plans), then the ncu pass (tools/profile_r2.sh).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2_pytest_gpu_1gpu.log 2>&1
python bench.py > gpurun_out/r2_bench_ddpg.json 2> gpurun_out/r2_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ddpg_steps20_warmup5.json 2>> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
for a in td3 sac tqc; do
  python bench.py --algo $a --no-cpu-baseline > gpurun_out/r2_bench_$a.json 2>> gpurun_out/r2_bench.err
done
OPRL_B200_CHAIN=1 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_ddpg_plan_full_chain.json 2>> gpurun_out/r2_bench.err
OPRL_B200_CHAIN_ACTOR=1 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_ddpg_plan_actor_chain.json 2>> gpurun_out/r2_bench.err
bash tools/profile_r2.sh r2 > gpurun_out/r2_profile.log 2>&1
tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "value %.0f us/step %.1f e2e %.0f launches %s frac %s cpu %s" % (
            d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d.get("launches_per_update"), r.get("frac"),
            (d.get("cpu_baseline") or {}).get("value")))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
cat gpurun_out/r2_pytest_gpu_1gpu.log
ls -la gpurun_out | grep r2_
