# Round-2 state-of-the-repo pass on one B200: full GPU suite, full bench line, launch lists, bandwidth-kernel ncu rows.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=12 2>&1 | grep -v "^E  \|^    " | tail -60 > gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras"
timeout 300 $CMD > gpurun_out/r2_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/r2_launches_b8_heun_step.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu list rc=$?"
bash tools/r2_bw_capture.sh
