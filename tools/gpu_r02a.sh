#!/bin/bash
# Round-2 first visit: smoke, the new parity tests first, the whole GPU suite, microbench (PDL / early loads / fp32 split), bench.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.csv 2>&1
nproc > gpurun_out/${TAG}_nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== new parity tests"; timeout 900 python -m pytest tests/test_parity_elementwise_gpu.py -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_new.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.txt
echo "== microbench sd21 bf16"; timeout 300 python tools/microbench.py --workload sd21 --dtypes bf16 --prompts 1 8 --variants mma-red mma-red-early mma-red-nopdl 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sd21 fp32"; timeout 300 python tools/microbench.py --workload sd21 --dtypes fp32 --prompts 1 8 --variants mma-red mma-red-early simt-red 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sd15"; timeout 300 python tools/microbench.py --workload sd15 --dtypes fp16 fp32 --prompts 1 --variants mma-red mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== microbench sdxl70"; timeout 300 python tools/microbench.py --workload sdxl70 --dtypes fp16 --prompts 1 2 --variants mma-red mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -12
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 6000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
ls -la gpurun_out | tail -8
