#!/bin/bash
# round 2, call W: graded tail of the work items: parity tests, scaling table, tail-min sweep, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2w_tests.txt
python tools/br_scaling.py 1024 2048 4096 8192 > gpurun_out/r2w_scaling.txt 2>&1
for tm in 100000 26 13 6 3; do echo "== TAIL_MIN $tm" >> gpurun_out/r2w_scaling.txt; TFHE_B200_BR_TAIL_MIN=$tm python tools/br_scaling.py 1024 4096 >> gpurun_out/r2w_scaling.txt 2>&1; done
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2w_bench.txt 2>&1
cat gpurun_out/r2w_tests.txt gpurun_out/r2w_scaling.txt; tail -1 gpurun_out/r2w_bench.txt | cut -c1-200
