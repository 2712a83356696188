#!/bin/bash
# v4 layer pipeline (N parts / progressive K issue): parity first (short timeout: a protocol mistake is a hang), then timeline + bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== umma bench"; timeout 120 python tests/prof_umma.py 2>&1 | tee gpurun_out/h_umma.log | tail -12
echo "== parity (tc)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shuffle.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/h_pytest.log
for parts in 3 2 1; do
  echo "== parts $parts"
  B200PETS_TC_PARTS=$parts timeout 200 python tests/prof_timeline.py 2>&1 | tee gpurun_out/h_timeline_p$parts.log | head -3 | cut -c1-400
  grep "coarse CTA 0: kernel" gpurun_out/h_timeline_p$parts.log | cut -c1-200
  B200PETS_TC_PARTS=$parts timeout 200 python tests/prof_rollout.py 2>&1 | tail -1
  B200PETS_TC_PARTS=$parts timeout 200 python tests/prof_rollout.py 32 2>&1 | tail -1
done
echo "== full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/h_pytest_all.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-scan 2>&1 | tail -1 | tee gpurun_out/h_bench.json | cut -c1-900
