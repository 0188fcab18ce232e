mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_c5_multi.py > gpurun_out/c24_c5_n8.txt 2>&1; tail -2 gpurun_out/c24_c5_n8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/c24_bench_n4.txt 2>&1; tail -1 gpurun_out/c24_bench_n4.txt | cut -c1-300
