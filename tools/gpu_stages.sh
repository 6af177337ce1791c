#!/bin/bash
mkdir -p gpurun_out
for a in tqc sac td3 ddpg; do
  OPRL_B200_DUMP_STAGES=1 timeout 300 python tools/stage_profile.py --algo $a > gpurun_out/r2_stage_costs_$a.txt 2> gpurun_out/r2_stage_plan_$a.txt
  tail -3 gpurun_out/r2_stage_costs_$a.txt
done
