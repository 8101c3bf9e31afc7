set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp12.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3 >> $L
echo "== conv bench" >> $L
timeout 200 python tools/conv_bench.py 8 2>&1 | awk '/^B8/ {print}' >> $L
FTB_CONV_DBG=1 timeout 200 python tools/conv_bench.py 8 2>&1 | awk '/issuer0/ {print}' | head -3 >> $L
timeout 400 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_bench12.json 2>> $L; echo "bench rc=$?" >> $L
tail -20 $L
