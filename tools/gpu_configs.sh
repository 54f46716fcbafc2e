#!/bin/bash
# BASELINE configs 3-5 on one GPU (the multi-GPU forms shard these per rank): bench lines for the record.
mkdir -p gpurun_out
for cfg in "sdxl 1" "sd21 8" "sdxl 2"; do
  set -- $cfg
  echo "== workload $1 prompts $2"
  timeout 900 python bench.py --workload $1 --prompts $2 --skip-cpu --steps 30 > gpurun_out/r01_cfg_$1_p$2.json 2> gpurun_out/r01_cfg_$1_p$2.err
  python - "$1" "$2" <<'PY'
import json, sys
w, p = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f'gpurun_out/r01_cfg_{w}_p{p}.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'dtype')}, 'frac', round(d['roofline']['frac'], 4), 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'eager', d['e2e']['eager_ms_per_step'], d['hook_overhead'])
except Exception as e:
    print('FAILED', e); print(open(f'gpurun_out/r01_cfg_{w}_p{p}.err').read()[-1500:])
PY
done
