#!/bin/bash
# round-2 evidence pass: full suite, bench (both arms), ncu launch list + full capture of the committed rollout kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/g_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/g_smoke.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/g_bench_ref.json
echo "== bench ours"; timeout 1200 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/g_bench.json
echo "== timeline"; timeout 300 python tests/prof_timeline.py 2>&1 | tee gpurun_out/g_timeline.log | head -4
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-scan > gpurun_out/g_launches_stdout.log 2>&1
tail -2 gpurun_out/g_launches_stdout.log | cut -c1-200
echo "== ncu full (config 2, pop 500)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc_kernel -s 6 -c 1 -f -o gpurun_out/g_rollout_tc_pop500 python tests/prof_rollout.py > gpurun_out/g_ncu_full_pop500.log 2>&1
tail -2 gpurun_out/g_ncu_full_pop500.log
echo "== ncu full (pop 16000)"
timeout 900 ncu --set full --clock-control none -k regex:rollout_tc_kernel -s 6 -c 1 -f -o gpurun_out/g_rollout_tc_pop16000 python tests/prof_rollout.py 32 > gpurun_out/g_ncu_full_pop16000.log 2>&1
tail -2 gpurun_out/g_ncu_full_pop16000.log
ls -la gpurun_out | tail -12
