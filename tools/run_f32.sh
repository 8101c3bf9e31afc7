#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fp32_mode" 2>&1 | tail -40 > gpurun_out/pytest_f32.log
tail -30 gpurun_out/pytest_f32.log
FTB_F32_KEEP=1 timeout 300 python tools/f32_diag.py > gpurun_out/f32_diag.log 2>&1
tail -60 gpurun_out/f32_diag.log
