#!/bin/bash
mkdir -p gpurun_out
for d in 208 192; do
echo "== debug $d"
(OPRL_B200_CHAIN_DEBUG=$d timeout 100 python tools/chain_prof.py ddpg 2>&1 | grep "chain 0\|chain 1\|op  [0-8]\|chain launches\|rror" | cut -c1-150 | tail -20)
done > gpurun_out/c2_debug.log 2>&1
cat gpurun_out/c2_debug.log
