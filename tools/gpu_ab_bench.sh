#!/bin/bash
# A/B under bench.py conditions (L2 flushed between steps): default library vs every go-tfhe_b200/lib/exp_*.so, twice, alternating
mkdir -p gpurun_out
OUT=gpurun_out/${1:-ab}_bench_ab.txt
: > $OUT
for rep in 1 2; do
  for so in default go-tfhe_b200/lib/exp_*.so; do
    [ "$so" = default ] || [ -e "$so" ] || continue
    if [ "$so" = default ]; then unset TFHE_B200_LIB; else export TFHE_B200_LIB=$PWD/$so; fi
    echo "== $so (rep $rep)" >> $OUT
    timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f  ms_per_step %.3f  kernel_ms %.3f  ks_ms %.3f  e2e %.0f  single %.3f' % (d['value'], d['ms_per_step'], d['stage_ms']['blind_rotate'], d['stage_ms']['key_switch'], d['e2e']['value'], d['single_gate_ms']))" >> $OUT 2>&1
  done
done
unset TFHE_B200_LIB
cat $OUT
