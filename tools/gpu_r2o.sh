#!/bin/bash
# round 2, call O: full GPU suite + full bench line after the tiled key switch / N = 2048 prefetch change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2o_tests.txt
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2o_bench.txt 2>&1
cat gpurun_out/r2o_tests.txt; tail -5 gpurun_out/r2o_bench.txt | cut -c1-400
