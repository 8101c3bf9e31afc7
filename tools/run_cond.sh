#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "wgrad or dgrad or cond" 2>&1 | tail -60 > gpurun_out/pytest_cond.log
tail -40 gpurun_out/pytest_cond.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fp32_mode" 2>&1 | tail -15 > gpurun_out/pytest_f32.log
tail -8 gpurun_out/pytest_f32.log
