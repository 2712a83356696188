#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== expectation cases"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "expectation" 2>&1 | tail -8 | tee gpurun_out/o_pytest_exp.log
echo "== full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/o_pytest.log
timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
