#!/bin/bash
# round 2, call Z: final ncu captures of the kernels that changed late (tiled key switch v4, N = 2048 blind rotation, N = 1024 blind rotation)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_tile_kernel -s 1 -c 1 -o gpurun_out/r02_ks_tile_uint5 -f python tools/pbs_run.py uint5 2048 2 tile > gpurun_out/r2z_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/r02_br_n2048 -f python tools/pbs_run.py uint5 2048 2 > gpurun_out/r2z_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/r02_br -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2z_ncu3.log 2>&1
ls -la gpurun_out/r02_ks_tile_uint5.ncu-rep gpurun_out/r02_br_n2048.ncu-rep gpurun_out/r02_br.ncu-rep
