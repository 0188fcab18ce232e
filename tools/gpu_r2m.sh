#!/bin/bash
# round 2, call M: tiled key switch v2 (templated, block order for L2 sharing): tests, timings, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "key_switch or pbs_uint" 2>&1 | tail -5 > gpurun_out/r2m_tests.txt
for v in gather tile; do for cnt in 256 2048 8192; do timeout 300 python tools/pbs_run.py uint5 $cnt 3 $v; done; done > gpurun_out/r2m_uint5_ks.txt 2>&1
for v in gather tile; do timeout 300 python tools/pbs_run.py uint2 4096 3 $v; timeout 300 python tools/pbs_run.py uint3 2048 3 $v; timeout 300 python tools/pbs_run.py uint4 2048 3 $v; done > gpurun_out/r2m_uint234_ks.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_tile_kernel -s 1 -c 1 -o gpurun_out/r02_ks_tile_uint5 -f python tools/pbs_run.py uint5 2048 2 tile > gpurun_out/r2m_ncu.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/pbs_run.py uint2 300 1 tile 2>&1 | tail -3 > gpurun_out/r2m_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python tools/pbs_run.py uint2 300 1 tile 2>&1 | tail -3 >> gpurun_out/r2m_memcheck.txt
cat gpurun_out/r2m_memcheck.txt gpurun_out/r2m_tests.txt gpurun_out/r2m_uint5_ks.txt gpurun_out/r2m_uint234_ks.txt
