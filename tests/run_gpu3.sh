#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/t_all.log
echo "== scale"; for s in 1 2 8 32; do timeout 120 python tests/prof_rollout.py $s 2>&1 | tail -1; done | tee gpurun_out/scale.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 4 -c 1 -o gpurun_out/prof_r1_v2 python tests/prof_rollout.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
