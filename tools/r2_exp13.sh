set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp13.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3 >> $L
echo "== wgrad bench" >> $L
timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
FTB_WGRAD_DBG=1 timeout 200 python tools/wgrad_bench.py 8 2>&1 | grep "dbg issuer" | awk 'NR%3==1' | head -12 >> $L
timeout 400 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_bench13.json 2>> $L; echo "bench rc=$?" >> $L
tail -30 $L
