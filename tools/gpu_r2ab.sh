#!/bin/bash
# ptxas --register-usage-level sweep: headline bench + Uint5 / Uint3 PBS + single-gate latency per library
mkdir -p gpurun_out
OUT=gpurun_out/r2ab_rul.txt
: > $OUT
for so in default go-tfhe_b200/lib/exp_*.so default; do
  if [ "$so" = default ]; then unset TFHE_B200_LIB; else export TFHE_B200_LIB=$PWD/$so; fi
  echo "== $so" >> $OUT
  python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f  kernel_ms %.3f  e2e %.0f  single %.3f' % (d['value'], d['stage_ms']['blind_rotate'], d['e2e']['value'], d['single_gate_ms']))" >> $OUT 2>&1
  python tools/pbs_run.py uint5 2048 3 2>&1 | tail -1 | sed -e 's/blind_rotate_launches.*key_switch_ms/ks/' >> $OUT
  python tools/pbs_run.py uint3 2048 3 2>&1 | tail -1 | sed -e 's/blind_rotate_launches.*key_switch_ms/ks/' >> $OUT
  python tools/pbs_run.py uint5 1 5 2>&1 | tail -1 | sed -e 's/blind_rotate_launches.*key_switch_ms/ks/' >> $OUT
done
unset TFHE_B200_LIB
cat $OUT
