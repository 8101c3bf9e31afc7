# wgrad experiments: MN-major MMA probe, tile shape / item order / L2 promotion sweeps, correctness of each variant
set -u
mkdir -p gpurun_out
L=gpurun_out/r2_wgrad_exp.log
: > $L
timeout 120 tools/mma_probe_mn >> $L 2>&1
for v in "" "FTB_WGRAD_TW=16" "FTB_WGRAD_TW=32" "FTB_WGRAD_INTERLEAVE=1" "FTB_WGRAD_TW=32 FTB_WGRAD_INTERLEAVE=1"; do
  echo "== correctness [$v]" >> $L
  env $v timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "wgrad" 2>&1 | tail -3 >> $L
done
for v in "" "FTB_TMA_PROMO=3" "FTB_TMA_PROMO=0" "FTB_WGRAD_INTERLEAVE=1" "FTB_WGRAD_INTERLEAVE=1 FTB_TMA_PROMO=3" \
         "FTB_WGRAD_TW=16" "FTB_WGRAD_TW=32" "FTB_WGRAD_TW=32 FTB_WGRAD_INTERLEAVE=1" "FTB_WGRAD_TW=32 FTB_WGRAD_INTERLEAVE=1 FTB_TMA_PROMO=3" \
         "FTB_WGRAD_TW=16 FTB_WGRAD_INTERLEAVE=1" "FTB_WGRAD_NOPAIR=1"; do
  echo "== bench [$v]" >> $L
  env $v timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
done
echo "== bench B=2 (L2-resident operands)" >> $L
timeout 200 python tools/wgrad_bench.py 2 2>&1 | tail -9 >> $L
for v in "" "FTB_TMA_PROMO=3" "FTB_TMA_PROMO=0"; do
  echo "== conv bench [$v]" >> $L
  env $v timeout 200 python tools/conv_bench.py 8 2>&1 | grep -v "^\[" | tail -12 >> $L
done
echo "== conv bench B=2" >> $L
timeout 200 python tools/conv_bench.py 2 2>&1 | tail -12 >> $L
tail -5 $L
