#!/bin/bash
# round 2, call U (8 GPUs): the bench under torchrun at N=8 (weak-scaled headline, strong-scaled c5, one-call multi-device c5)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2u_gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2u_bench_n8.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 --no-configs ) > gpurun_out/r2u_bench_n4.txt 2>&1
wc -l gpurun_out/r2u_gpus.txt; tail -4 gpurun_out/r2u_bench_n8.txt | cut -c1-300; tail -4 gpurun_out/r2u_bench_n4.txt | cut -c1-200
