#!/bin/bash
# round 2, call B: same-box A/B of key-load policy / carried accumulator words / round-1 kernel + ncu capture of the default
mkdir -p gpurun_out
{
echo "== default lib"; timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
for so in go-tfhe_b200/lib/exp_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"
  TFHE_B200_LIB=$PWD/$so timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
done
echo "== default lib again"; timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -1
} > gpurun_out/r2b_variants.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/r2b_br -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b_variants.txt
