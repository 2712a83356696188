#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/u_pytest.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-scan 2>&1 | tail -1 | tee gpurun_out/u_bench.json | cut -c1-260
