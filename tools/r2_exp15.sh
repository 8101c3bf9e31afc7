set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp15.log
: > $L
for v in "FTB_WGRAD_MAP4=1" "FTB_WGRAD_MAP4=1 FTB_WGRAD_TW=16" ""; do
  echo "== [$v]" >> $L
  env $v timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k wgrad 2>&1 | tail -1 >> $L
  env $v timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
done
tail -40 $L
