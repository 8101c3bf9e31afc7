#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unet3d" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_0.json 2> gpurun_out/ab_0.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_0.json')); print(d['value'], d['ms_per_step'], d['train']['ms_per_step'], d['gpu_launches'])"
