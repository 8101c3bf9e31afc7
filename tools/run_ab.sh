#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1 2; do
  unset FTB_NO_NSPLIT FTB_NO_NZ_SPREAD
  if [ $v = 1 ]; then export FTB_NO_NZ_SPREAD=1; fi
  if [ $v = 2 ]; then export FTB_NO_NSPLIT=1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "variant=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print(d['value'], d['ms_per_step'], d['train']['ms_per_step'], d['gpu_launches'])"
done
