#!/bin/bash
# round 2, call P: paired transforms for the L = 1 sets (exp lib) against the default
mkdir -p gpurun_out
OUT=gpurun_out/r2p_pair.txt
: > $OUT
for so in default go-tfhe_b200/lib/exp_*.so default; do
  if [ "$so" = default ]; then unset TFHE_B200_LIB; else export TFHE_B200_LIB=$PWD/$so; fi
  echo "== $so" >> $OUT
  for cfg in "uint5 2048" "uint4 2048" "uint3 2048" "uint2 4096"; do timeout 300 python tools/pbs_run.py $cfg 3 >> $OUT 2>&1; done
done
for so in go-tfhe_b200/lib/exp_*.so; do
  echo "== tests with $so" >> $OUT
  TFHE_B200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f4.py -m gpu -x -q -k "uint or pbs or many_lut or indexed" 2>&1 | tail -3 >> $OUT
done
unset TFHE_B200_LIB
sed -e 's/blind_rotate_launches.*key_switch_ms/ks/' $OUT
