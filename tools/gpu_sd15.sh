#!/bin/bash
mkdir -p gpurun_out
echo "== trace tests"; timeout 900 python -m pytest tests/test_trace_gpu.py -q 2>&1 | tail -4
echo "== bench sd15"; timeout 900 python bench.py --workload sd15 --skip-cpu > gpurun_out/r01_bench_sd15.json 2> gpurun_out/r01_bench_sd15.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r01_bench_sd15.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'dtype')}, 'frac', round(d['roofline']['frac'], 4), 'e2e', d['e2e']['value'], d['hook_overhead'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r01_bench_sd15.err').read()[-1500:])
PY
echo "== microbench sd15 variants"; timeout 600 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import microbench
sys.argv = ['microbench', '--workload', 'sd15', '--dtypes', 'fp32', 'fp16', '--prompts', '1', '--variants', 'mma-red', 'simt-red']
microbench.main()
PY
