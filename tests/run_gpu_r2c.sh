#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/c_pytest.log
for c in 9 10 11 8; do
  echo "== timeline H0CHUNKS=$c"; B200PETS_TC_H0CHUNKS=$c timeout 300 python tests/prof_timeline.py 2>&1 | grep -v "^coarse CTA 40" | tee gpurun_out/c_timeline_c$c.log
done
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -3 | tee gpurun_out/c_bench.log
