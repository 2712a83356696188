#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/f_pytest.log
echo "== timeline"; timeout 300 python tests/prof_timeline.py 2>&1 | grep -v "^coarse CTA 40" | tee gpurun_out/f_timeline.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -2 | tee gpurun_out/f_bench.log
