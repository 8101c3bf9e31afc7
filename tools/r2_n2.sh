set -u
mkdir -p gpurun_out
L=gpurun_out/r2_n2.log
: > $L
nvidia-smi -L >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "nccl" 2>&1 | tail -4 >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2>> $L; echo "bench n2 rc=$?" >> $L
tail -12 $L
