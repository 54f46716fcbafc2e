#!/bin/bash
TAG=${1:-r01m}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1700 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== ncu finalize + accumulate in an e2e run"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize -c 4 -f -o gpurun_out/${TAG}_prof_finalize \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu > gpurun_out/${TAG}_ncu_fin.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_fin.log | cut -c1-200
