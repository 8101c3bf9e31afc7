set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp4.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 >> $L
echo "== wgrad bench" >> $L
timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
echo "== conv bench" >> $L
timeout 200 python tools/conv_bench.py 8 2>&1 | grep -v "^\[" | tail -12 >> $L
timeout 400 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_bench4.json 2>> $L; echo "bench rc=$?" >> $L
tail -30 $L
