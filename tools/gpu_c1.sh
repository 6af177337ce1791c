#!/bin/bash
# chain-kernel bring-up: per-tensor gradient errors vs the oracle, chain profile, for accumulator group sizes 2 / 4 / 8
mkdir -p gpurun_out
for g in 2 4 8; do
  export OPRL_B200_CHAIN_GROUP=$g
  (timeout 120 python tools/debug_parity.py ddpg 2>&1 | grep -v "target" | head -30) > gpurun_out/c1_dbg_ddpg_g$g.log
  (timeout 120 python tools/debug_parity.py td3 2>&1 | grep "grad\|loss" | head -30) > gpurun_out/c1_dbg_td3_g$g.log
  (timeout 200 python tools/chain_prof.py ddpg 2>&1 | tail -24) > gpurun_out/c1_prof_ddpg_g$g.log
  (timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "ddpg or td3" -s 2>&1 | grep "param L2\|passed\|failed" ) > gpurun_out/c1_parity_g$g.log
done
tail -n 30 gpurun_out/c1_prof_ddpg_g2.log gpurun_out/c1_prof_ddpg_g4.log gpurun_out/c1_parity_g*.log
