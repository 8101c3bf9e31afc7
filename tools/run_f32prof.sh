#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/f32_one.py > gpurun_out/f32_one.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/f32_launches.csv python tools/f32_one.py > gpurun_out/f32_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/f32_one.log
