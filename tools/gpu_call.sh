mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree" > gpurun_out/c29_pytest.txt 2>&1; tail -3 gpurun_out/c29_pytest.txt
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize.py --variant=mg > gpurun_out/c29_race_mg.txt 2>&1; grep -v "^=========     " gpurun_out/c29_race_mg.txt | tail -4
TFHE_B200_BR=mg timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -2
