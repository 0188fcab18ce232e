#!/bin/bash
# round 2, call N: N = 2048 blind rotation (Uint5, 2048 PBS): work-item length sweep and build variants
mkdir -p gpurun_out
OUT=gpurun_out/r2n_uint5_br.txt
: > $OUT
for st in 0 36 54 72 83 108 134 179 357 1071; do
  echo "== chunk steps $st" >> $OUT
  TFHE_B200_BR_CHUNK_STEPS=$st timeout 300 python tools/pbs_run.py uint5 2048 3 >> $OUT 2>&1
done
for so in default go-tfhe_b200/lib/exp_*.so default; do
  if [ "$so" = default ]; then unset TFHE_B200_LIB; else export TFHE_B200_LIB=$PWD/$so; fi
  echo "== $so" >> $OUT
  timeout 300 python tools/pbs_run.py uint5 2048 3 >> $OUT 2>&1
done
unset TFHE_B200_LIB
sed -e 's/blind_rotate_launches.*key_switch_ms/ks/' $OUT
