#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== full GPU suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/m_pytest.log
for h0 in 0 9; do
  echo "== h0 chunks $h0 (0 = default 8)"
  B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_timeline.py 2>&1 > gpurun_out/m_timeline_$h0.log; head -1 gpurun_out/m_timeline_$h0.log | cut -c1-400
  grep "^mma layer [012]" gpurun_out/m_timeline_$h0.log
  grep "coarse CTA 0: kernel" gpurun_out/m_timeline_$h0.log | cut -c1-160
  B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_H0CHUNKS=$h0 timeout 200 python tests/prof_rollout.py 32 2>&1 | tail -1
done
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-scan 2>&1 | tail -1 | tee gpurun_out/m_bench.json | cut -c1-300
