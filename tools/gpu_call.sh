mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/c30_pytest.txt 2>&1; tail -14 gpurun_out/c30_pytest.txt
