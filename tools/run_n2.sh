#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q -s -k "data_parallel" 2>&1 | tail -8 > gpurun_out/ddp_test.log
cat gpurun_out/ddp_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
