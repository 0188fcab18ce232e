mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c12_pytest.txt 2>&1; tail -3 gpurun_out/c12_pytest.txt
timeout 900 python tools/bench_configs.py c4 > gpurun_out/c12_configs.txt 2>&1; cat gpurun_out/c12_configs.txt
