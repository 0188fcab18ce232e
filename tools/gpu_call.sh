set -x
mkdir -p gpurun_out
./tools/fp64_bench > gpurun_out/c3_fp64.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "key_switch" > gpurun_out/c3_pytest_ks.txt 2>&1
( VARIANTS="ldg" bash tools/exp_variants.sh ) > gpurun_out/c3_variants.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.txt 2>&1
tail -15 gpurun_out/c3_pytest_ks.txt; tail -3 gpurun_out/c3_pytest.txt; grep "^==\|^BR\|^iter 2" gpurun_out/c3_variants.txt; cat gpurun_out/c3_fp64.txt
