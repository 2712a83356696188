#!/bin/bash
# multi-GPU pass (gpurun --gpus N): NCCL equality test + bench.py at N ranks (weak scaling + config 5 strong scaling)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee gpurun_out/e_gpus.log
echo "== multi-GPU NCCL test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/e_pytest_multi.log
echo "== bench 1 GPU"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-scan 2>&1 | tail -1 | tee gpurun_out/e_bench_n1.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench $n GPUs, NCCL all-gather"
    B200PETS_PEER_EXCHANGE=0 NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-scan5 2>&1 | grep '^{' | tail -1 | tee gpurun_out/e_bench_n${n}_nccl.log | cut -c1-330
    echo "== bench $n GPUs"
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 2>&1 | grep '^{' | tail -1 | tee gpurun_out/e_bench_n$n.log
  fi
done
