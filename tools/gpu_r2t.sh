#!/bin/bash
# round 2, call T: CUDA-graph replay of the circuit runner: tests, sanitizer on a small circuit, bench c3 with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_parity.py -m gpu -x -q -k "circuit or adder" 2>&1 | tail -6 > gpurun_out/r2t_tests.txt
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2t_bench.txt 2>&1
cat gpurun_out/r2t_tests.txt; tail -4 gpurun_out/r2t_bench.txt | cut -c1-200
