#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity (tc)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shuffle.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/j_pytest.log
for spin in 1 0; do
  echo "== spin $spin"
  B200PETS_TC_SPIN=$spin timeout 200 python tests/prof_timeline.py 2>&1 | tee gpurun_out/j_timeline_s$spin.log | head -1 | cut -c1-400
  grep -A1 "^mma layer [012]" gpurun_out/j_timeline_s$spin.log
  grep "coarse CTA 0: kernel" gpurun_out/j_timeline_s$spin.log | cut -c1-160
  B200PETS_TC_SPIN=$spin timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_SPIN=$spin timeout 200 python tests/prof_rollout.py 32 2>&1 | tail -1
done
