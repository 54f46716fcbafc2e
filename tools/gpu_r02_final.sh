#!/bin/bash
# Round-2 closing visit: smoke, the whole GPU suite, finalize microbenchmark + capture, the bench lines kept under profiles/.
TAG=${1:-r02z}
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== hook breakdown"; timeout 300 python tools/profile_hook.py 2>&1 | tail -32
echo "== finalize"; timeout 300 python tools/microbench_finalize.py --workload sd21 2>&1 | tail -1
timeout 300 python tools/microbench_finalize.py --workload sdxl70 2>&1 | tail -1
echo "== ncu finalize"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize_fast -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_finalize \
  python tools/microbench_finalize.py --no-save --workload sd21 > gpurun_out/${TAG}_ncu_fin.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_fin.log | cut -c1-200
echo "== microbench"; timeout 300 python tools/microbench.py --workload sd21 --dtypes bf16 fp32 --prompts 1 8 --variants mma-red-early mma-red mma-red-nopdl 2>&1 | grep -v "per_layer\": true" | tail -12
timeout 300 python tools/microbench.py --workload sd15 --dtypes fp16 fp32 --prompts 1 8 --variants mma-red-early 2>&1 | grep -v "per_layer\": true" | tail -4
B="timeout 900 python bench.py --warmup 5"
echo "== bench (config 2)"; $B --steps 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== bench fp32"; $B --steps 20 --dtype fp32 --skip-cpu --skip-eager > gpurun_out/${TAG}_bench_fp32.json 2> gpurun_out/${TAG}_bench_fp32.err; tail -c 900 gpurun_out/${TAG}_bench_fp32.json; tail -3 gpurun_out/${TAG}_bench_fp32.err
echo "== bench sd15 fp32"; $B --steps 20 --workload sd15 --skip-cpu --skip-eager > gpurun_out/${TAG}_bench_sd15.json 2> gpurun_out/${TAG}_bench_sd15.err; tail -c 900 gpurun_out/${TAG}_bench_sd15.json; tail -3 gpurun_out/${TAG}_bench_sd15.err
echo "== reference arm (host cores)"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-400 gpurun_out/${TAG}_bench_ref.json; tail -2 gpurun_out/${TAG}_bench_ref.err
