#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t_all.log
echo "== timeline"; timeout 200 python tests/prof_timeline.py 2>&1 | grep -E "epilogue|fine" | tee gpurun_out/timeline.log
echo "== scale"; for s in 1 8 32; do timeout 120 python tests/prof_rollout.py $s 2>&1 | tail -1; done | tee gpurun_out/scale.log
