#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list + one full capture of the dominant kernel.
# usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tests|bench|ncu|all] [batch for ncu]
set -u
what=${1:-all}
nb=${2:-8}
mkdir -p gpurun_out
if [[ $what == all || $what == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [[ $what == all || $what == ncu ]]; then
  CMD="python bench.py --steps 1 --warmup 1 --batch $nb --no-e2e --no-cpu-baseline --no-train --no-extras"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 4 -c 5 \
      -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
