mkdir -p gpurun_out
for v in tma tex w16 tmem tmex tmex+tma; do echo "== $v"; timeout 600 compute-sanitizer --tool synccheck python tools/sanitize.py --variant=$v 2>&1 | grep -v "^=========     " | grep "=========\|done" | head -5; done > gpurun_out/c18_synccheck.txt 2>&1
cat gpurun_out/c18_synccheck.txt
