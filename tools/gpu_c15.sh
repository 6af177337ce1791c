#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "odd_batch" 2>&1 | grep "B=\|passed\|failed\|rror\|skip" | cut -c1-160
