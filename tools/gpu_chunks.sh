#!/bin/bash
echo "== accumulate tests"; timeout 900 python -m pytest tests/test_accumulate_gpu.py -q -x 2>&1 | tail -6
echo "== microbench"; timeout 600 python tools/microbench.py --workload sd21 --dtypes bf16 fp32 --prompts 1 --variants mma-red 2>&1 | grep -v "per_layer\": true" | tail -3
