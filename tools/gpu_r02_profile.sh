#!/bin/bash
# Round-2 profile visit: suite, chunked-head-dim and finalize microbenchmarks, the ncu launch list of the bench command,
# full captures of the accumulate kernel (16-bit form in the bench command, fp32 split form in the microbenchmark) and the
# DRAM traffic per launch, isolated (L2 flushed) and in steady state (--cache-control none over the rotating sets).
TAG=${1:-r02f}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== microbench"; timeout 300 python tools/microbench.py --workload sd15 --dtypes fp16 fp32 --prompts 1 8 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -4
timeout 300 python tools/microbench.py --workload sd21 --dtypes bf16 fp32 --prompts 1 --variants mma-red-early mma-red mma-red-nopdl 2>&1 | grep -v "per_layer\": true" | tail -6
echo "== finalize"; timeout 300 python tools/microbench_finalize.py --workload sd21 2>&1 | tail -1
timeout 300 python tools/microbench_finalize.py --workload sdxl70 2>&1 | tail -1
echo "== ncu launch list (the bench command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu --skip-eager > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200
echo "== ncu full: accumulate, 16-bit form (bench command)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_mma -s 10 -c 2 -f -o gpurun_out/${TAG}_prof_mma \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu --skip-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
echo "== ncu full: accumulate, fp32 split form"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_mma -s 30 -c 1 -f -o gpurun_out/${TAG}_prof_fp32 \
  python tools/microbench.py --no-save --workload sd21 --dtypes fp32 --prompts 1 --variants mma-red-early > gpurun_out/${TAG}_ncu_fp32.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_fp32.log | cut -c1-200
echo "== DRAM traffic per launch: steady state (no cache control) and isolated"
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:accumulate_mma -s 60 -c 10 --csv"
timeout 600 ncu $M --cache-control none --log-file gpurun_out/${TAG}_traffic_steady.csv python tools/microbench.py --no-save --workload sd21 --dtypes bf16 --prompts 1 --variants mma-red-early > /dev/null 2>&1
timeout 600 ncu $M --log-file gpurun_out/${TAG}_traffic_isolated.csv python tools/microbench.py --no-save --workload sd21 --dtypes bf16 --prompts 1 --variants mma-red-early > /dev/null 2>&1
tail -4 gpurun_out/${TAG}_traffic_steady.csv | cut -c1-300
ls -la gpurun_out | grep ${TAG}
