#!/bin/bash
# cache-policy experiments for the key loads: s* = N <= 1024 policy (128-bit + Uint3), b* = N = 2048 policy (Uint5 + Uint4)
mkdir -p gpurun_out
OUT=gpurun_out/r2ah_policy.txt
: > $OUT
for so in default go-tfhe_b200/lib/exp_s*.so default; do
  if [ "$so" = default ]; then unset TFHE_B200_LIB; else export TFHE_B200_LIB=$PWD/$so; fi
  echo "== $so" >> $OUT
  python tools/pbs_run.py 128 4096 3 2>&1 | tail -1 >> $OUT
  python tools/pbs_run.py uint3 2048 3 2>&1 | tail -1 >> $OUT
done
for so in default go-tfhe_b200/lib/exp_b*.so default; do
  if [ "$so" = default ]; then unset TFHE_B200_LIB; else export TFHE_B200_LIB=$PWD/$so; fi
  echo "== $so" >> $OUT
  python tools/pbs_run.py uint5 2048 3 2>&1 | tail -1 >> $OUT
  python tools/pbs_run.py uint4 2048 3 2>&1 | tail -1 >> $OUT
done
unset TFHE_B200_LIB
sed -e 's/blind_rotate_launches.*key_switch_ms/ks/' $OUT
