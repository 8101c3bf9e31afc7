# last check of the round: GPU suite, smoke(), default bench line
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2v_pytest.log; tail -2 gpurun_out/r2v_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r2v_smoke.log
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2v_bench_ref.json | cut -c1-400
