set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp6.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3 >> $L
for v in "" "FTB_NISS=1" "FTB_CONV_NO_FAST27=1"; do
  echo "== conv bench [$v]" >> $L
  env $v timeout 200 python tools/conv_bench.py 8 2>&1 | grep -v "^\[" | tail -9 >> $L
done
echo "== wgrad bench" >> $L
timeout 200 python tools/wgrad_bench.py 8 2>&1 | tail -9 >> $L
timeout 400 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_bench6.json 2>> $L; echo "bench rc=$?" >> $L
FTB_NISS=1 timeout 400 python bench.py --no-cpu-baseline --no-extras --no-train > gpurun_out/r2_bench6_niss1.json 2>> $L; echo "bench niss1 rc=$?" >> $L
tail -45 $L
