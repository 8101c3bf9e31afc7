set -u
mkdir -p gpurun_out
L=gpurun_out/r2_exp17.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $L
for v in "" "FTB_TMA_NARROW=1"; do
echo "== conv bench [$v]" >> $L
env $v timeout 200 python tools/conv_bench.py 8 2>&1 | awk '/^B8/ {print}' >> $L
done
timeout 400 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_bench17.json 2>> $L; echo "bench rc=$?" >> $L
FTB_TMA_NARROW=1 timeout 400 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_bench17_narrow.json 2>> $L; echo "bench rc=$?" >> $L
tail -30 $L
