#!/bin/bash
# GPU-box driver script for development runs: separate processes so that a hang in one group cannot hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== umma selftest" ; timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "library or umma" -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/t_umma.log
echo "== f32 + optimisers" ; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not tc and not umma and not library and not in_kernel and not agent and not fused" -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/t_f32.log
echo "== tc" ; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tc or in_kernel or agent or fused" -p no:cacheprovider 2>&1 | tail -80 | tee gpurun_out/t_tc.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -5 | tee gpurun_out/bench.log
