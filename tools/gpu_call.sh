mkdir -p gpurun_out
( VARIANTS="" bash tools/exp_variants.sh ) > gpurun_out/c31_variants.txt 2>&1
grep "^==\|^BR\|correct" gpurun_out/c31_variants.txt | cut -c1-200
