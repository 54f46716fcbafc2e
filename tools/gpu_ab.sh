#!/bin/bash
echo "== accumulate tests"; timeout 600 python -m pytest tests/test_accumulate_gpu.py -q -x 2>&1 | tail -3
echo "== microbench sd21"; timeout 600 python tools/microbench.py --workload sd21 --dtypes bf16 --variants mma-red mma-red-static 2>&1 | grep -v "per_layer\": true" | tail -8
echo "== microbench sdxl"; timeout 600 python tools/microbench.py --workload sdxl --dtypes fp16 --prompts 1 --variants mma-red mma-red-static 2>&1 | grep -v "per_layer\": true" | tail -8
