set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp14.log
: > $L
for v in "" "FTB_WGRAD_TW=16" "FTB_WGRAD_TW=32"; do
  echo "== [$v]" >> $L
  env $v timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k wgrad 2>&1 | tail -1 >> $L
  env $v timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
  env $v FTB_WGRAD_DBG=1 timeout 200 python tools/wgrad_bench.py 8 2>&1 | grep "dbg issuer" | awk 'NR%3==1' | head -4 >> $L
done
tail -50 $L
