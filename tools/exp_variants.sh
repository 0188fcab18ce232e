#!/bin/bash
# A/B of the three key-fetch variants of the blind-rotate kernel (+ any experimental engine builds)
for v in ldg tma tex; do
  echo "== TFHE_B200_BR=$v"; TFHE_B200_BR=$v python tools/gpu_quick.py 128 4096 2>&1 | tail -2
done
for so in go-tfhe_b200/lib/exp_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"
  TFHE_B200_LIB=$PWD/$so python tools/gpu_quick.py 128 4096 2>&1 | tail -2
done
