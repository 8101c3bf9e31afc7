#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 141 -c 14 -f -o gpurun_out/prof_k1up $CMD > gpurun_out/ncu_k1up.log 2>&1
echo rc=$?
