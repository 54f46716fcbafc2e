#!/bin/bash
TAG=${1:-r01i}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== microbench sd21"; timeout 600 python tools/microbench.py --workload sd21 --dtypes bf16 --variants mma-red mma-red-static 2>&1 | grep -v "per_layer\": true" | tail -8
echo "== microbench sdxl"; timeout 600 python tools/microbench.py --workload sdxl --dtypes fp16 --prompts 1 2 --variants mma-red mma-red-static 2>&1 | grep -v "per_layer\": true" | tail -8
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/'+ "$TAG" + '_bench.json').read().strip().splitlines()[-1]) if False else None
PY
tail -c 1500 gpurun_out/${TAG}_bench.json | head -c 1500; tail -3 gpurun_out/${TAG}_bench.err
