#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/t_all.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
echo "== timeline"; timeout 200 python tests/prof_timeline.py 2>&1 | grep -E "epilogue|fine|mma layer" | tee gpurun_out/timeline.log
