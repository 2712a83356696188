#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== peer exchange kernels on one GPU"
timeout 300 python -m pytest tests/test_gpu_peer_exchange.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s_pytest_peer.log
echo "== full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/s_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/s_bench.json | cut -c1-200
