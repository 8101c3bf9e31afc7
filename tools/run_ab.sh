#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 4 5 8; do
  export FTB_K1_NZ=$v
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "K1_NZ=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print(d['value'], d['ms_per_step'], d['roofline']['conv1x1']['ms'], d['train']['ms_per_step'])"
done
