#!/bin/bash
# round 2, call S: compute-sanitizer over the round-2 paths + launch list of a bench run
mkdir -p gpurun_out
OUT=gpurun_out/r02_compute_sanitizer.txt
: > $OUT
timeout 300 python tools/sanitize_r2.py > gpurun_out/r2s_plain.txt 2>&1; tail -2 gpurun_out/r2s_plain.txt >> $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_r2.py" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_r2.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done|Error|error|hazard" | head -20 >> $OUT
done

cat $OUT
