# end-of-round check of HEAD + a launch list of the 128^3 SDE sampler (configs[4])
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r2v2_bench.json 2> gpurun_out/r2v2_bench.err; echo "bench rc=$?"
timeout 300 python tools/sde128_bench.py 3 > gpurun_out/r2v2_sde128_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
   --log-file gpurun_out/r2v2_sde128_launches.csv python tools/sde128_bench.py 3 > gpurun_out/r2v2_ncu_sde.log 2>&1
echo "ncu sde128 rc=$?"
