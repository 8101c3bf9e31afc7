#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "trilinear or grads" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --train-only > gpurun_out/ab_t.json 2> gpurun_out/ab_t.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_t.json')); print(d['value'], d['ms_per_step'])"
