set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp7.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv or unet3d" 2>&1 | tail -3 >> $L
for v in "" "FTB_WSLOT=4" "FTB_WSLOT=5" "FTB_WSLOT=6" "FTB_WSLOT=8"; do
  echo "== conv bench [$v]" >> $L
  env $v FTB_CONV_DBG=1 FTB_CONV_PLAN=1 timeout 200 python tools/conv_bench.py 8 2>&1 | awk '/^conv plan/ && !seen[$0]++ {print} /^B8|issuer0/ {print}' >> $L
done
tail -90 $L
