#!/bin/bash
mkdir -p gpurun_out
for d in 0 2 4 6; do
  echo "=== debug $d"
  OPRL_B200_CHAIN_GROUP=4 OPRL_B200_CHAIN_DEBUG=$d timeout 200 python tools/chain_prof.py ddpg 2>&1 | grep -v "first arrivals\|read-out done\|MMA warp" | tail -34
done > gpurun_out/c2_debug.log 2>&1
cat gpurun_out/c2_debug.log
