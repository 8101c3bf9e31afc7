#!/bin/bash
set -u
mkdir -p gpurun_out
export FTB_BENCH_MINIMAL=1
CMD="python bench.py --train-only --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/train_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 860 -c 900 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/ncu_train.log 2>&1
echo rc=$?; cat gpurun_out/train_plain.log | tail -1
