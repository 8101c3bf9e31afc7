#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q -s -k "nccl" 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/ensemble_bench.py 64 8 100 euler > gpurun_out/ensemble_n2.json 2> gpurun_out/ensemble_n2.err
echo "ens n2 rc=$?"; tail -2 gpurun_out/ensemble_n2.json; tail -3 gpurun_out/ensemble_n2.err
