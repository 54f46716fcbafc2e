#!/bin/bash
mkdir -p gpurun_out
echo "== baseline"; timeout 200 python tools/microbench.py --no-save --workload sd21 --dtypes bf16 --prompts 1 8 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== tile-major accumulator addressing (timing experiment)"; DAAM_EXP_TILE_MAJOR=1 timeout 200 python tools/microbench.py --no-save --workload sd21 --dtypes bf16 --prompts 1 8 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== sdxl70 baseline / tile-major"; timeout 200 python tools/microbench.py --no-save --workload sdxl70 --dtypes fp16 --prompts 1 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -1
DAAM_EXP_TILE_MAJOR=1 timeout 200 python tools/microbench.py --no-save --workload sdxl70 --dtypes fp16 --prompts 1 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -1
echo "== sd21_768"; timeout 600 python bench.py --workload sd21_768 --steps 20 --warmup 5 --skip-cpu --skip-eager > gpurun_out/r02_bench_sd21_768.json 2> gpurun_out/r02_bench_sd21_768.err; tail -c 1200 gpurun_out/r02_bench_sd21_768.json; tail -3 gpurun_out/r02_bench_sd21_768.err
