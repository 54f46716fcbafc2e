#!/bin/bash
# 8-GPU visit (torchrun, one rank per GPU, NCCL): BASELINE configs 4 and 5 as written, the default config for scaling,
# and the NCCL gather test. Usage: gpurun --gpus 8 -- 'bash tools/gpu_r02_n8.sh [N]'
N=${1:-8}
TAG=r02_n${N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpus.csv 2>&1
nproc > gpurun_out/${TAG}_nproc.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== NCCL gather test"; timeout 600 python -m pytest tests/test_distributed_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_dist.txt
echo "== cfg2 default (scaling line)"; timeout 900 $RUN --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --skip-eager > gpurun_out/${TAG}_cfg2.json 2> gpurun_out/${TAG}_cfg2.err; tail -c 2500 gpurun_out/${TAG}_cfg2.json; tail -3 gpurun_out/${TAG}_cfg2.err
echo "== cfg4: SD-2.1, 8 prompts per GPU, 50 steps"; timeout 900 $RUN --master-port 29512 bench.py --gpus $N --workload sd21 --prompts 8 --steps 50 --warmup 5 --skip-eager > gpurun_out/${TAG}_cfg4.json 2> gpurun_out/${TAG}_cfg4.err; tail -c 2500 gpurun_out/${TAG}_cfg4.json; tail -3 gpurun_out/${TAG}_cfg4.err
echo "== cfg5: SDXL, 70 layers, 2 prompts per GPU, 30 steps"; timeout 1200 $RUN --master-port 29513 bench.py --gpus $N --workload sdxl70 --prompts 2 --steps 30 --warmup 5 --skip-eager > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err; tail -c 2500 gpurun_out/${TAG}_cfg5.json; tail -3 gpurun_out/${TAG}_cfg5.err
ls -la gpurun_out | grep ${TAG}
