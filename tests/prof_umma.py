"""Diagnostics: tcgen05.mma issue/completion rate per shared-memory operand layout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbrl_lib_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
for mode in (0, 3, 4):
    for (k, n) in ((208, 208), (256, 256), (208, 144), (208, 112), (208, 80), (208, 64), (208, 48), (208, 32), (208, 16), (64, 208)):
        for reps in (8,):
            _lib.check(lib.b200pets_debug_umma_bench(mode, k, n, reps, _lib.ptr(out), _lib.stream_ptr()))
            torch.cuda.synchronize()
            iss, tot = out.cpu().tolist()
            nm = reps * (k // 16)
            print(f"mode {mode} k {k} n {n}: {nm} MMAs, issue {iss/nm:.1f} cyc/MMA, complete {tot/nm:.1f} cyc/MMA (floor {128*n/256:.0f})")
