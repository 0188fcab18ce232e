#!/bin/bash
# round 2, call A: GPU tests + A/B of work-item sizes and accumulator layouts (one box, one call)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt
{
for cs in 700 0 100 70 35 25; do
  echo "== default lib, chunk_steps=$cs"; TFHE_B200_BR_CHUNK_STEPS=$cs timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -2
done
for so in go-tfhe_b200/lib/exp_*.so; do
  [ -e "$so" ] || continue
  for cs in 700 0; do
    echo "== $so chunk_steps=$cs"
    TFHE_B200_BR_CHUNK_STEPS=$cs TFHE_B200_LIB=$PWD/$so timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -2
  done
done
} > gpurun_out/r2a_variants.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.txt 2>&1
tail -3 gpurun_out/r2a_pytest.txt; cat gpurun_out/r2a_variants.txt | grep -E "==|BR gates"
