#!/bin/bash
# 1-GPU twins of BASELINE configs 3, 4, 5 (+ fp32 config-1 shapes and SD-1.5), bench lines kept under gpurun_out/r02_cfg*_n1.json
mkdir -p gpurun_out
B="timeout 900 python bench.py --warmup 5 --skip-eager"
echo "== cfg3 sdxl fp16, 1 prompt, 50 steps"; $B --workload sdxl --steps 50 > gpurun_out/r02_cfg3_n1.json 2> gpurun_out/r02_cfg3_n1.err; tail -c 1800 gpurun_out/r02_cfg3_n1.json; tail -2 gpurun_out/r02_cfg3_n1.err
echo "== cfg4 twin: sd21, 8 prompts, 50 steps"; $B --workload sd21 --prompts 8 --steps 50 > gpurun_out/r02_cfg4_n1.json 2> gpurun_out/r02_cfg4_n1.err; tail -c 1800 gpurun_out/r02_cfg4_n1.json; tail -2 gpurun_out/r02_cfg4_n1.err
echo "== cfg5 twin: sdxl70, 2 prompts, 30 steps"; $B --workload sdxl70 --prompts 2 --steps 30 > gpurun_out/r02_cfg5_n1.json 2> gpurun_out/r02_cfg5_n1.err; tail -c 1800 gpurun_out/r02_cfg5_n1.json; tail -2 gpurun_out/r02_cfg5_n1.err
echo "== fp32 sd21 (config 1 shapes)"; $B --workload sd21 --dtype fp32 --steps 20 --skip-cpu > gpurun_out/r02_sd21_fp32_n1.json 2> gpurun_out/r02_sd21_fp32_n1.err; tail -c 1500 gpurun_out/r02_sd21_fp32_n1.json; tail -2 gpurun_out/r02_sd21_fp32_n1.err
echo "== sd15 fp32"; $B --workload sd15 --steps 20 --skip-cpu > gpurun_out/r02_sd15_fp32_n1.json 2> gpurun_out/r02_sd15_fp32_n1.err; tail -c 1500 gpurun_out/r02_sd15_fp32_n1.json; tail -2 gpurun_out/r02_sd15_fp32_n1.err
