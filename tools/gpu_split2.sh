#!/bin/bash
echo "== accumulate tests"; timeout 600 python -m pytest tests/test_accumulate_gpu.py -q -x 2>&1 | tail -3
echo "== microbench fp32"; timeout 600 python tools/microbench.py --workload sd21 --dtypes fp32 bf16 --variants mma-red 2>&1 | grep -v "per_layer\": true" | tail -6
