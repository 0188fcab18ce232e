#!/bin/bash
# round 2, call J (2 GPUs): the whole GPU suite with both devices visible (multi-device context on a real peer), then the
# bench under torchrun at N=2 (NCCL key broadcast, strong-scaled c5, one-call multi-device c5) and the tools C5 runner.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2j_tests.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2j_bench_n2.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > gpurun_out/r2j_ref_n2.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/bench_c5_multi.py --log2 18 > gpurun_out/r2j_c5_tool.txt 2>&1
cat gpurun_out/r2j_gpus.txt gpurun_out/r2j_tests.txt; tail -5 gpurun_out/r2j_bench_n2.txt | cut -c1-1500; tail -3 gpurun_out/r2j_ref_n2.txt | cut -c1-400; tail -2 gpurun_out/r2j_c5_tool.txt | cut -c1-500
