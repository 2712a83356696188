#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity (tc)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shuffle.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/l_pytest.log
for h0 in 0 8 10; do
  echo "== h0 chunks $h0 (0 = default 9)"
  B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_timeline.py 2>&1 > gpurun_out/l_timeline_$h0.log; head -1 gpurun_out/l_timeline_$h0.log | cut -c1-400
  grep "^mma layer [012]" gpurun_out/l_timeline_$h0.log
  grep "coarse CTA 0: kernel" gpurun_out/l_timeline_$h0.log | cut -c1-160
  B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 32 2>&1 | tail -1
done
