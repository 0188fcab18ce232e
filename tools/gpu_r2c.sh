#!/bin/bash
# round 2, call C: gates-per-block A/B (L1 sharing through the per-step block barrier) + full GPU tests on the default build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.txt
{
echo "== default lib"; timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
for so in go-tfhe_b200/lib/exp_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"
  TFHE_B200_LIB=$PWD/$so timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
done
echo "== default lib again"; timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
echo "== default lib configs c4 c5 c3"; timeout 600 python tools/bench_configs.py c4 c5 c3 2>&1 | tail -4
} > gpurun_out/r2c_variants.txt 2>&1
tail -3 gpurun_out/r2c_pytest.txt; cat gpurun_out/r2c_variants.txt
