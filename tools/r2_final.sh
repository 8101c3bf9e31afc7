# Round-2 evidence pass on one B200: GPU suite, full bench line, launch lists (sampling step + training step),
# ncu --set full rows of the conv / wgrad / attention kernels, bandwidth-kernel rows.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 2>&1 | grep -v "^E  \|^    " | tail -30 > gpurun_out/r2f_pytest_gpu.log
tail -3 gpurun_out/r2f_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-train --no-extras"
timeout 300 $CMD > gpurun_out/r2f_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/r2f_launches_b8_heun_step.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 2 \
   -f -o /tmp/r2f_prof_conv $CMD > gpurun_out/r2f_ncu_conv.log 2>&1
echo "ncu conv rc=$?"
python tools/ncu_summary.py /tmp/r2f_prof_conv.ncu-rep > gpurun_out/r2f_conv_igemm_ncu_full.csv
timeout 600 ncu --set full --clock-control none -k regex:'kvctx|qout' -s 0 -c 2 \
   -f -o /tmp/r2f_prof_attn $CMD > gpurun_out/r2f_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
python tools/ncu_summary.py /tmp/r2f_prof_attn.ncu-rep > gpurun_out/r2f_attention_ncu_full.csv
TCMD="python bench.py --train-only --steps 1 --warmup 1 --no-cpu-baseline"
FTB_BENCH_MINIMAL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r2f_train_step_launches.csv $TCMD > gpurun_out/r2f_ncu_train.log 2>&1
echo "ncu train list rc=$?"
FTB_BENCH_MINIMAL=1 timeout 900 ncu --set full --clock-control none -k regex:'wgrad_kernel' -s 130 -c 2 \
   -f -o /tmp/r2f_prof_wgrad $TCMD > gpurun_out/r2f_ncu_wgrad.log 2>&1
echo "ncu wgrad rc=$?"
python tools/ncu_summary.py /tmp/r2f_prof_wgrad.ncu-rep > gpurun_out/r2f_wgrad_ncu_full.csv
FTB_BENCH_MINIMAL=1 timeout 900 ncu --set full --clock-control none -k regex:'normact_bwd' -s 20 -c 3 \
   -f -o /tmp/r2f_prof_nab $TCMD > gpurun_out/r2f_ncu_nab.log 2>&1
echo "ncu normact_bwd rc=$?"
python tools/ncu_summary.py /tmp/r2f_prof_nab.ncu-rep > gpurun_out/r2f_normact_bwd_ncu_full.csv
