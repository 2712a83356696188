#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t_all.log
echo "== bench"; timeout 900 python bench.py --steps 20 --no-cpu 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench.log
