#!/bin/bash
# round 2, call E: keep-own exchanges A/B, new host API tests + GPU tests + bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_pytest.txt
{
echo "== default lib"; timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
for so in go-tfhe_b200/lib/exp_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"
  TFHE_B200_LIB=$PWD/$so timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
done
echo "== default lib again"; timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
} > gpurun_out/r2e_variants.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.txt 2>&1
tail -3 gpurun_out/r2e_pytest.txt; cat gpurun_out/r2e_variants.txt; cut -c1-400 gpurun_out/r2e_bench.txt
