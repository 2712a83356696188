#!/bin/bash
# round-2 GPU validation pass A: new shuffle / parity tests on the v3 pipeline (KR off), then the K-round pipeline
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== KR=0 full GPU suite"; 
B200PETS_TC_KR=0 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/a_pytest_kr0.log
echo "== KR=1 tensor-core tests";
B200PETS_TC_KR=1 timeout 900 python -m pytest tests -m gpu -q -k "tc or closed_form or shuffle or fused or plan or agent or mpc or sharded or deterministic or mbpo or step" 2>&1 | tail -25 | tee gpurun_out/a_pytest_kr1.log
for kr in 0 1; do
  echo "== timeline KR=$kr"; B200PETS_TC_KR=$kr timeout 300 python tests/prof_timeline.py 2>&1 | tee gpurun_out/a_timeline_kr$kr.log
  echo "== bench KR=$kr"; B200PETS_TC_KR=$kr timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/a_bench_kr$kr.log
done
