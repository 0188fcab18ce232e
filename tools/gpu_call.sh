mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.txt 2>&1; tail -2 gpurun_out/c6_pytest.txt
timeout 600 python bench.py > gpurun_out/c6_bench.txt 2>&1; tail -1 gpurun_out/c6_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c6_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_mma_kernel -s 1 -c 1 -o gpurun_out/c6_ks_mma -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c6_ncu_ks.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/c6_br -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c6_ncu_br.log 2>&1
ls -la gpurun_out
