#!/bin/bash
# End-of-round evidence with the final code: tests, smoke, bench, reference arm, ncu launch list + full capture.
TAG=${1:-r01z}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 500 gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu --skip-eager > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-160
echo "== ncu full (accumulate_mma, value leg)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_mma -s 3 -c 2 -f -o gpurun_out/${TAG}_prof_mma \
  python bench.py --steps 5 --warmup 3 --skip-overhead --skip-cpu --skip-eager > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log | cut -c1-160
