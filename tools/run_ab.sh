#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "trilinear or unet3d or grads" 2>&1 | tail -4
for v in 0 1; do
  if [ $v = 1 ]; then export FTB_TRILINEAR_DIRECT=1; else unset FTB_TRILINEAR_DIRECT; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "DIRECT=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print(d['value'], d['ms_per_step'], d['train']['ms_per_step'])"
done
