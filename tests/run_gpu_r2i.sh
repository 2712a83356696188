#!/bin/bash
# v4 layer pipeline, rounds decoupled from parts + per-warp arrivals: parity, timeline, A/B over the first part's width
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity (tc)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shuffle.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/i_pytest.log
for part0 in 8 13 4; do
  echo "== part0 $part0 chunks"
  B200PETS_TC_PART0=$part0 timeout 200 python tests/prof_timeline.py 2>&1 | tee gpurun_out/i_timeline_p$part0.log | head -1 | cut -c1-400
  grep -A1 "^mma layer [12]" gpurun_out/i_timeline_p$part0.log
  grep "coarse CTA 0: kernel" gpurun_out/i_timeline_p$part0.log | cut -c1-160
  B200PETS_TC_PART0=$part0 timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_PART0=$part0 timeout 200 python tests/prof_rollout.py 32 2>&1 | tail -1
done
