set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "adaptive" 2>&1 | grep -v "^E  \|^    " | tail -40 > gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_adaptive_solvers_vs_oracle 2>&1 | tail -15 > gpurun_out/r2_t4.log
tail -4 gpurun_out/r2_t4.log
for cfg in "- - -" "1 - -" "- 4 -" "- 5 -" "1 5 -" "- 6 -"; do
  set -- $cfg
  env $( [ "$1" != "-" ] && echo FTB_NISS=$1 ) $( [ "$2" != "-" ] && echo FTB_WSLOT=$2 ) FTB_CONV_DBG=1 timeout 120 python tools/conv_bench.py 8 2>&1 | sed "s/^/[niss=$1 wslot=$2] /" >> gpurun_out/r2_conv_sweep.log
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
