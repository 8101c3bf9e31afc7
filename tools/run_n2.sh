#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "nccl" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; wc -l gpurun_out/bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['train']['value'], d['train']['ms_per_step'], d['train']['e2e']['ms_per_step'], d['clocks'])"
