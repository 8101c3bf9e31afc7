#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 1 2 3 4; do
  export FTB_KV_SPLIT=$v
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-train > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "KV_SPLIT=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print(d['value'], d['ms_per_step'])"
done
