#!/bin/bash
# quick development check: tensor-core parity subset, one-step timeline, rollout-only scaling
mkdir -p gpurun_out
echo "== tc tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tc or in_kernel or agent or fused or step" -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/t_tc.log
echo "== timeline"; timeout 200 python tests/prof_timeline.py 2>&1 | grep -E "epilogue|fine" | tee gpurun_out/timeline.log
echo "== scale"; for s in 1 32; do timeout 120 python tests/prof_rollout.py $s 2>&1 | tail -1; done | tee gpurun_out/scale.log
