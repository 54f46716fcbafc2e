#!/bin/bash
# Round-2 second visit: reduce-mode variants (segmented / per-row bulk reduces), the re-tiled finalize, fused word expansion.
TAG=${1:-r02b}
mkdir -p gpurun_out
echo "== tests (new kernels)"; timeout 900 python -m pytest tests/test_parity_elementwise_gpu.py tests/test_finalize_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_new.txt
echo "== microbench sd21 bf16"; timeout 300 python tools/microbench.py --workload sd21 --dtypes bf16 --prompts 1 8 --variants mma-red-early mma-seg-early mma-rows-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sd21 fp32"; timeout 300 python tools/microbench.py --workload sd21 --dtypes fp32 --prompts 1 8 --variants mma-red-early mma-seg-early mma-rows-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sd15"; timeout 300 python tools/microbench.py --workload sd15 --dtypes fp16 fp32 --prompts 1 --variants mma-red-early mma-seg-early mma-rows-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sdxl70"; timeout 300 python tools/microbench.py --workload sdxl70 --dtypes fp16 --prompts 1 --variants mma-red-early mma-seg-early mma-rows-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== finalize"; timeout 300 python tools/microbench_finalize.py --workload sd21 2>&1 | tail -3
timeout 300 python tools/microbench_finalize.py --workload sdxl70 2>&1 | tail -3
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.txt
