#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite (PDL on)"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/d_pytest.log
echo "== GPU suite TC subset (PDL off)"
B200PETS_PDL=0 timeout 900 python -m pytest tests -m gpu -q -k "tc or shuffle or fused or plan or closed_form" 2>&1 | tail -5 | tee gpurun_out/d_pytest_nopdl.log
echo "== timeline"; timeout 300 python tests/prof_timeline.py 2>&1 | grep -v "^coarse CTA 40" | tee gpurun_out/d_timeline.log
echo "== bench PDL on"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -2 | tee gpurun_out/d_bench.log
echo "== bench PDL off"; B200PETS_PDL=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-scan 2>&1 | tail -2 | tee gpurun_out/d_bench_nopdl.log
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_shuffle.py -q -k "tile_shuffle_matches_oracle or mbpo_device or shard_invariance" 2>&1 | tail -8 | tee gpurun_out/d_memcheck.log
