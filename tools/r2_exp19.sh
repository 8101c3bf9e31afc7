set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp19.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -q -x 2>&1 | tail -2 >> $L
for v in "" "FTB_NO_WRING_GROW=1"; do
echo "== [$v]" >> $L
env $v FTB_CONV_PLAN=1 timeout 200 python tools/conv_bench.py 8 2>&1 | awk '/^conv plan/ && !seen[$0]++ {print} /^B8/ {print}' | tail -10 >> $L
env $v timeout 400 python bench.py --no-cpu-baseline --no-extras --no-train > gpurun_out/r2_bench19_$([ -z "$v" ] && echo grow || echo nogrow).json 2>> $L; echo "bench rc=$?" >> $L
done
tail -30 $L
