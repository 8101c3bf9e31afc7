set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp16.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3 >> $L
timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
timeout 400 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_bench16.json 2>> $L; echo "bench rc=$?" >> $L
tail -20 $L
