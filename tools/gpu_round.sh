#!/bin/bash
# One GPU-box visit: smoke, GPU tests, bench, micro-benchmark, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root on the box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv 2>&1
nproc > gpurun_out/${TAG}_nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== microbench"; timeout 600 python tools/microbench.py --workload sd21 2>&1 | tail -40
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
echo "== ncu full (accumulate)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_acc \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
ls -la gpurun_out | tail -15
