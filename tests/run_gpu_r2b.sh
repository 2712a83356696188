#!/bin/bash
# round-2 GPU validation pass B: full suite on the v3.5 kernel (output epilogue rewrite, 9/4 N split, fused refit kernel)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/b_pytest.log
for c in 9 7 10; do
  echo "== timeline H0CHUNKS=$c"; B200PETS_TC_H0CHUNKS=$c timeout 300 python tests/prof_timeline.py 2>&1 | tee gpurun_out/b_timeline_c$c.log
done
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/b_bench.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -2 | tee gpurun_out/b_bench_ref.log
