#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "cond_flow_trainer or flow_trainer_step" 2>&1 | tail -60 > gpurun_out/pytest_cond.log
tail -40 gpurun_out/pytest_cond.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "conditioning or ensemble or decode" 2>&1 | tail -40 > gpurun_out/pytest_front.log
tail -30 gpurun_out/pytest_front.log
