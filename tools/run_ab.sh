#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unet3d or trilinear or conv" 2>&1 | tail -5 > gpurun_out/pytest_ab.log
tail -5 gpurun_out/pytest_ab.log
for v in 0 1; do
  if [ $v = 1 ]; then export FTB_NO_PDL=1; else unset FTB_NO_PDL; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "NO_PDL=$v rc=$?"; cut -c1-400 gpurun_out/ab_$v.json; tail -2 gpurun_out/ab_$v.err
done
