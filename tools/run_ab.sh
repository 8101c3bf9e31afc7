#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do
  if [ $v = 1 ]; then export FTB_TRAIN_NO_ALIAS=1; else unset FTB_TRAIN_NO_ALIAS; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --train-only > gpurun_out/ab_t$v.json 2> gpurun_out/ab_t$v.err
  echo "NO_ALIAS=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_t$v.json')); print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
done
