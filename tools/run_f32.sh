#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fp32_mode" 2>&1 | tail -40 > gpurun_out/pytest_f32.log
tail -12 gpurun_out/pytest_f32.log
FTB_F32_KEEP=1 timeout 300 python tools/f32_diag.py 2>&1 | tail -4 > gpurun_out/f32_diag.log
cat gpurun_out/f32_diag.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/f32_launches.csv python tools/f32_one.py > gpurun_out/f32_ncu.log 2>&1
