#!/bin/bash
# round 2, call I: ncu captures of the other shipped kernels (N = 2048 throughput kernel, Uint5 gather key switch, both latency
# kernels) + observed Uint tolerances (pytest -s) + the full bench line
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/r02_br_n2048 -f python tools/pbs_run.py uint5 2048 2 > gpurun_out/r02_ncu_n2048.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:key_switch_kernel -s 1 -c 1 -o gpurun_out/r02_ks_gather_uint5 -f python tools/pbs_run.py uint5 2048 2 > gpurun_out/r02_ncu_ksg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_lat_kernel -s 2 -c 1 -o gpurun_out/r02_br_lat -f python tools/one_gate.py ldg > gpurun_out/r02_ncu_lat.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_latp_kernel -s 1 -c 1 -o gpurun_out/r02_br_latp -f python tools/pbs_run.py uint5 1 2 > gpurun_out/r02_ncu_latp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "pbs_uint_sets" 2>&1 | grep -E "max .delta|passed|failed" > gpurun_out/r02_uint_tolerances.txt
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2i_bench.txt 2>&1
ls -la gpurun_out/r02_* | head -20; cat gpurun_out/r02_uint_tolerances.txt; tail -4 gpurun_out/r2i_bench.txt | cut -c1-600
