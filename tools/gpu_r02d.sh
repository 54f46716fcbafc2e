#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_finalize_gpu.py tests/test_trace_gpu.py tests/test_parity_elementwise_gpu.py -q -x 2>&1 | tail -6
echo "== finalize"; timeout 300 python tools/microbench_finalize.py --workload sd21 2>&1 | tail -1
timeout 300 python tools/microbench_finalize.py --workload sdxl70 2>&1 | tail -1
MB="python tools/microbench.py --workload sd21 --dtypes fp32 --prompts 1 8 --variants mma-red-early"
echo "== fp32 baseline"; timeout 200 $MB 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== fp32 8 converter warps"; DAAM_SPLIT_CONV_WARPS=8 timeout 200 $MB 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== fp32 no hi rewrite"; DAAM_SPLIT_NO_HI=1 timeout 200 $MB 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== fp32 both"; DAAM_SPLIT_NO_HI=1 DAAM_SPLIT_CONV_WARPS=8 timeout 200 $MB 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== bf16 one CTA per SM"; DAAM_MMA_ONE_CTA=1 timeout 200 python tools/microbench.py --workload sd21 --dtypes bf16 --prompts 1 8 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -2
echo "== no-hi parity"; DAAM_SPLIT_NO_HI=1 timeout 300 python -m pytest tests/test_parity_elementwise_gpu.py -q -k "tf32 or (seeded and float32)" 2>&1 | tail -8
echo "== ncu fp32"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_mma -s 30 -c 1 -f -o gpurun_out/${TAG}_prof_fp32 \
  python tools/microbench.py --workload sd21 --dtypes fp32 --prompts 1 --variants mma-red-early > gpurun_out/${TAG}_ncu_fp32.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_fp32.log | cut -c1-200
echo "== reference arm on the GPU (torch eager)"; timeout 600 python bench.py --impl reference --ref-device cuda --steps 20 --warmup 5 > gpurun_out/${TAG}_ref_cuda.json 2> gpurun_out/${TAG}_ref_cuda.err; cat gpurun_out/${TAG}_ref_cuda.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_ref_cuda.err
