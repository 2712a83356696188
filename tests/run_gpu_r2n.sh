#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity (tc)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shuffle.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/n_pytest.log
for ts in 0 1; do
 for h0 in 0 9; do
  echo "== tail split $ts, h0 chunks $h0 (0 = default 8)"
  B200PETS_TC_TAILSPLIT=$ts B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_TAILSPLIT=$ts B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_TAILSPLIT=$ts B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 32 2>&1 | tail -1
 done
done
