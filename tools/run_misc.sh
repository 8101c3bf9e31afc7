#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_split.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo rc=$?
FTB_CONV_PLAN=1 timeout 300 python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras 2>&1 | grep "conv plan" | sort | uniq -c | sort -rn | head -60 > gpurun_out/conv_plans.txt
