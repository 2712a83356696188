#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_all.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launch_bench.log 2>&1
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 4 -c 1 -o gpurun_out/prof_r1_final python tests/prof_rollout.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_after.txt
