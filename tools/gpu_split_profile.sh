#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_accumulate_gpu.py tests/test_trace_gpu.py -q 2>&1 | tail -3
echo "== ncu full: fp32 split form"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_mma -s 3 -c 1 -f -o gpurun_out/r01_prof_mma_fp32 \
  python bench.py --dtype fp32 --steps 5 --warmup 3 --skip-overhead --skip-cpu --skip-eager > gpurun_out/r01_ncu_fp32.log 2>&1
tail -1 gpurun_out/r01_ncu_fp32.log | cut -c1-160
