#!/bin/bash
# chain-kernel bring-up: per-tensor gradient errors vs the oracle, then a short bench
mkdir -p gpurun_out
(timeout 120 python tools/debug_parity.py ddpg 2>&1 | head -40) > gpurun_out/c1_dbg_ddpg.log
(timeout 120 python tools/debug_parity.py td3 2>&1 | head -50) > gpurun_out/c1_dbg_td3.log
(timeout 200 python bench.py --no-cpu-baseline --steps 500 2>&1 | tail -3) > gpurun_out/c1_bench.log
head -c 3000 gpurun_out/c1_dbg_ddpg.log
tail -c 1500 gpurun_out/c1_bench.log
