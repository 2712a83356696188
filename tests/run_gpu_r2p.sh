#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== full GPU suite (merged refit + sample kernel)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/p_pytest.log
echo "== bench merged"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-scan 2>&1 | tail -1 | tee gpurun_out/p_bench_merged.json | cut -c1-260
echo "== bench 3 kernels"; B200PETS_CEM_MERGED=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-scan 2>&1 | tail -1 | tee gpurun_out/p_bench_split.json | cut -c1-260
echo "== launch list (merged)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-scan > gpurun_out/p_launches_stdout.log 2>&1
tail -1 gpurun_out/p_launches_stdout.log | cut -c1-100
