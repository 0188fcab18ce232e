#!/bin/bash
# round 2, call F: full bench.py (headline + configs block) at N=1, timing of the whole run
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2f_bench.txt 2>&1
tail -5 gpurun_out/r2f_bench.txt | cut -c1-3000
