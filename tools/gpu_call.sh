mkdir -p gpurun_out
( VARIANTS="ldg" bash tools/exp_variants.sh ) > gpurun_out/c21_variants.txt 2>&1
grep "^==\|^BR" gpurun_out/c21_variants.txt
