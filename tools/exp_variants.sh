#!/bin/bash
# A/B harness: default engine (optionally all key-fetch variants) + any experimental engine builds (lib/exp_*.so)
VARIANTS=${VARIANTS:-ldg}
for v in $VARIANTS; do
  echo "== default lib, TFHE_B200_BR=$v"; TFHE_B200_BR=$v timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -2
done
for so in go-tfhe_b200/lib/exp_*.so; do
  [ -e "$so" ] || continue
  echo "== $so (TFHE_B200_BR=${EXPBR:-ldg})"
  TFHE_B200_BR=${EXPBR:-ldg} TFHE_B200_LIB=$PWD/$so timeout 120 python tools/gpu_quick.py 128 4096 2>&1 | tail -2
done
