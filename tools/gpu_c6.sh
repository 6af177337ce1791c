#!/bin/bash
mkdir -p gpurun_out
cp oprl_b200/liboprl_b200.so /tmp/new.so
for v in old new; do
  if [ $v = old ]; then cp build/liboprl_old.so oprl_b200/liboprl_b200.so; else cp /tmp/new.so oprl_b200/liboprl_b200.so; fi
  for a in td3 ddpg; do
    timeout 300 python tools/stage_profile.py --algo $a --iters 1000 > gpurun_out/ab_stage_${a}_$v.txt 2>/dev/null
  done
done
cp /tmp/new.so oprl_b200/liboprl_b200.so
paste gpurun_out/ab_stage_td3_old.txt gpurun_out/ab_stage_td3_new.txt | cut -c1-150
paste gpurun_out/ab_stage_ddpg_old.txt gpurun_out/ab_stage_ddpg_new.txt | cut -c1-150
