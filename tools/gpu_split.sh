#!/bin/bash
echo "== accumulate tests"; timeout 600 python -m pytest tests/test_accumulate_gpu.py -q -x 2>&1 | tail -8
echo "== microbench fp32"; timeout 600 python tools/microbench.py --workload sd21 --dtypes fp32 --variants mma-red simt-red 2>&1 | grep -v "per_layer\": true" | tail -6
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
