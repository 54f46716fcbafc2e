#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_parity_elementwise_gpu.py -q -x -k "golden_layers" 2>&1 | tail -4
echo "== microbench sd21 bf16"; timeout 300 python tools/microbench.py --workload sd21 --dtypes bf16 --prompts 1 8 --variants mma-red-early mma-regs-early mma-ldst-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sd21 fp32"; timeout 300 python tools/microbench.py --workload sd21 --dtypes fp32 --prompts 1 8 --variants mma-red-early mma-regs-early mma-ldst-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sd15"; timeout 300 python tools/microbench.py --workload sd15 --dtypes fp16 fp32 --prompts 1 --variants mma-red-early mma-regs-early 2>&1 | grep -v "per_layer\": true" | tail -12
