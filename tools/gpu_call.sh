mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/c20_bench.txt 2>&1; tail -1 gpurun_out/c20_bench.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['gpu_launches'])"
