# ncu --set full rows of the HBM-bound kernels (tools/bw_kernels.py); the reports are summarised on the box
# (tools/ncu_summary.py) and deleted: only the CSVs travel back.
set -u
mkdir -p gpurun_out
timeout 300 python tools/bw_kernels.py > gpurun_out/r2_bw_plain.log 2>&1; echo "bw plain rc=$?"
FTB_BW_PART=standalone FTB_BW_NO_TIMING=1 timeout 900 ncu --profile-from-start off --set full --clock-control none \
   -k regex:'heun|axpy|rk4|interp_kernel|decode|trilinear|pack_unfold|normact|embed_kernel|drift|vote|lincomb|error_ratio|advance|sumsq' \
   -c 60 -f -o /tmp/prof_bw_a python tools/bw_kernels.py > gpurun_out/r2_ncu_bw_a.log 2>&1
echo "ncu bw standalone rc=$?"
python tools/ncu_summary.py /tmp/prof_bw_a.ncu-rep > gpurun_out/r2_bw_standalone_ncu_full.csv
FTB_BW_PART=train FTB_BW_NO_TIMING=1 timeout 900 ncu --profile-from-start off --set full --clock-control none \
   -k regex:'normact_bwd|trilinear_bwd|adam_kernel|ema_kernel|mse|chan_sum|sumsq_kernel|interp_kernel|embed_kernel' \
   -c 16 -f -o /tmp/prof_bw_b python tools/bw_kernels.py > gpurun_out/r2_ncu_bw_b.log 2>&1
echo "ncu bw train rc=$?"
python tools/ncu_summary.py /tmp/prof_bw_b.ncu-rep > gpurun_out/r2_bw_train_ncu_full.csv
ls -la /tmp/*.ncu-rep
