#!/bin/bash
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== microbench"; timeout 300 python tools/microbench.py --workload sd21 --dtypes bf16 fp32 --prompts 1 8 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -4
timeout 300 python tools/microbench.py --workload sd15 --dtypes fp16 fp32 --prompts 1 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== finalize"; timeout 300 python tools/microbench_finalize.py --workload sd21 2>&1 | tail -1
timeout 300 python tools/microbench_finalize.py --workload sdxl70 2>&1 | tail -1
echo "== ncu finalize"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize_fast -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_finalize \
  python tools/microbench_finalize.py --no-save --workload sd21 > gpurun_out/${TAG}_ncu_fin.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_fin.log | cut -c1-200
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 5000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== reference arm on the GPU (torch eager)"; timeout 600 python bench.py --impl reference --ref-device cuda --steps 20 --warmup 5 > gpurun_out/${TAG}_ref_cuda.json 2> gpurun_out/${TAG}_ref_cuda.err; cat gpurun_out/${TAG}_ref_cuda.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_ref_cuda.err
echo "== config twins"; bash tools/gpu_r02_cfg_n1.sh
