set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp18.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3 >> $L
timeout 400 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_bench18.json 2>> $L; echo "bench rc=$?" >> $L
FTB_BENCH_MINIMAL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r2_train_step_launches_c.csv python bench.py --train-only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_train_c.log 2>&1
echo "ncu train list rc=$?" >> $L
tail -8 $L
