#!/bin/bash
# Bring-up visit for the tcgen05 kernel: guarded tests first (a protocol bug traps after ~2 s instead of hanging).
TAG=${1:-r01b}
mkdir -p gpurun_out
echo "== accumulate tests"; timeout 600 python -m pytest tests/test_accumulate_gpu.py -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_acc.txt
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== microbench"; timeout 600 python tools/microbench.py --workload sd21 --dtypes bf16 2>&1 | tail -40
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
