#!/bin/bash
echo "== tests"; timeout 600 python -m pytest tests/test_accumulate_gpu.py tests/test_trace_gpu.py -q -x 2>&1 | tail -3
timeout 600 python tools/microbench.py --workload sd15 --dtypes fp32 --prompts 1 --variants mma-red 2>&1 | grep -v "per_layer\": true" | tail -1
timeout 600 python tools/microbench.py --workload sd21 --dtypes fp32 --prompts 1 8 --variants mma-red 2>&1 | grep -v "per_layer\": true" | tail -2
