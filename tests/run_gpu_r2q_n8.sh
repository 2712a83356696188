#!/bin/bash
# 8 GPUs: weak-scaling bench + config-5 scan with the peer-memory exchange (short: the GPU budget is charged x8)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=8
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep '^{' | tail -1 | tee gpurun_out/q_bench_n8_peer2.log | cut -c1-330
