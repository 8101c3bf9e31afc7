set -u
mkdir -p gpurun_out
L=gpurun_out/r2_wgrad2.log
: > $L
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -4 >> $L
echo "== bench partials" >> $L
timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
echo "== bench atomics" >> $L
FTB_WGRAD_ATOMIC=1 timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
echo "== bench partials B=2" >> $L
timeout 200 python tools/wgrad_bench.py 2 2>&1 | tail -9 >> $L
timeout 300 python bench.py --train-only --no-cpu-baseline > gpurun_out/r2_train_bench.json 2>> $L; echo "train bench rc=$?" >> $L
# conv kernel: full capture with source of two 64^3 launches (3^3 48->48 with the fused epilogue)
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 2 \
   -f -o gpurun_out/r2_prof_conv $CMD > gpurun_out/r2_ncu_conv.log 2>&1
echo "ncu conv rc=$?" >> $L
tail -30 $L
