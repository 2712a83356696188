#!/bin/bash
mkdir -p gpurun_out
echo "== scale tests"; timeout 900 python -m pytest tests/test_gpu_scale.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/t_scale.log
echo "== step timing"; timeout 300 python tests/prof_step.py 2>&1 | tail -3 | tee gpurun_out/step.log
