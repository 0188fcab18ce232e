#!/bin/bash
# ncu captures of the shipped kernels: full set + source for the throughput kernel (N=1024), launch list of a bench step
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/${TAG}_br -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_ncu_br.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_launches.log 2>&1
ls -la gpurun_out/${TAG}_*
