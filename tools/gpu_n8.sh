#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --skip-overhead --skip-cpu > gpurun_out/r01_scale_n$N.json 2> gpurun_out/r01_scale_n$N.err
tail -c 1500 gpurun_out/r01_scale_n$N.json; tail -3 gpurun_out/r01_scale_n$N.err | cut -c1-300
