set -u
mkdir -p gpurun_out
L=gpurun_out/r2_train3.log
: > $L
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -4 >> $L
timeout 300 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_train_bench_g3.json 2>> $L; echo "train bench (g3) rc=$?" >> $L
FTB_NORMACT_NO_G3=1 timeout 300 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_train_bench_nog3.json 2>> $L; echo "train bench (no g3) rc=$?" >> $L
FTB_WGRAD_ATOMIC=1 timeout 300 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_train_bench_atomic.json 2>> $L; echo "train bench (atomic wgrad) rc=$?" >> $L
FTB_BENCH_MINIMAL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r2_train_step_launches.csv python bench.py --train-only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_train.log 2>&1
echo "ncu train list rc=$?" >> $L
tail -12 $L
