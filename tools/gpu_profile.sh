#!/bin/bash
# Profile visit: tests, microbench (PDL on/off, SD-2.1 + SDXL), ncu launch list of the bench command, full capture of the accumulate kernel.
TAG=${1:-r01c}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== microbench sd21"; timeout 600 python tools/microbench.py --workload sd21 --dtypes bf16 --variants mma-red mma-red-nopdl simt-red 2>&1 | tail -20
echo "== microbench sdxl"; timeout 600 python tools/microbench.py --workload sdxl --dtypes fp16 --prompts 1 2 --variants mma-red mma-red-nopdl 2>&1 | tail -20
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2600 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== ncu launch list (the bench command, all kernels)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200
echo "== ncu full (accumulate_mma)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_mma -s 3 -c 2 -f -o gpurun_out/${TAG}_prof_mma \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ls -la gpurun_out | tail -8
