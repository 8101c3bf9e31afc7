#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kvctx|qout' -s 14 -c 2 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
