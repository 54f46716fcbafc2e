#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL gather test + torchrun bench at N and the reference arm under torchrun.
N=${1:-2}; TAG=${2:-r01f}
mkdir -p gpurun_out
nvidia-smi -L
echo "== distributed gpu test"; timeout 600 python -m pytest tests/test_distributed_gpu.py -q -x 2>&1 | tail -5
for n in 1 $N; do
  echo "== bench N=$n"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --skip-overhead --skip-cpu > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --skip-overhead --skip-cpu > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  fi
  tail -c 1800 gpurun_out/${TAG}_bench_n$n.json; tail -4 gpurun_out/${TAG}_bench_n$n.err
done
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 10 --warmup 2 2>/dev/null | tail -c 600
