#!/bin/bash
TAG=${1:-r01n}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 700 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== ncu finalize"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize -c 2 -f -o gpurun_out/${TAG}_prof_finalize \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu > gpurun_out/${TAG}_ncu_fin.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_fin.log | cut -c1-200
