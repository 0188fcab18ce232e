#!/bin/bash
# build an experimental engine: tools/build_exp.sh <tag> <extra nvcc flags...>  -> go-tfhe_b200/lib/exp_<tag>.so
# KERN=<mangled-name regex> selects which kernel's register/spill line is echoed (default: the LDG blind-rotate kernel)
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v -DTFHE_BR_SINGLE_TU "$@" \
  -o go-tfhe_b200/lib/exp_$tag.so go-tfhe_b200/csrc/tfhe_b200.cu 2> /tmp/exp_$tag.ptxas
grep -A2 "Function properties for.*${KERN:-blind_rotate_kernelILi10ELi3ELi6ELb1ELi[0-9]*ELb0}" /tmp/exp_$tag.ptxas | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $tag"
