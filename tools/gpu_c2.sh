#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2_pytest_gpu.log 2>&1
cat gpurun_out/r2_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3
