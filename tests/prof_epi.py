"""Diagnostics: hidden-layer epilogue pass (13 chunks x 128 rows, 16 warps) with / without concurrent tcgen05.mma."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbrl_lib_b200 import _lib
lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device="cuda:0")
reps = 20
def run(flags, nmma, n, label):
    out.zero_()
    _lib.check(lib.b200pets_debug_umma_bench(100 + flags, nmma, n, reps, _lib.ptr(out), _lib.stream_ptr()))
    torch.cuda.synchronize()
    e, m, mi, _ = out.cpu().tolist()
    print(f"{label:58s} epilogue {e/reps:7.0f} cyc/pass   MMAs: {m/max(nmma,1):6.1f} cyc each over {m} cyc (issue {mi})")
for pack, pl in ((0, "F2FP pack"), (32, "int half-up pack"), (64, "int RNE pack"), (128, "no pack")):
    run(0 | pack, 0, 208, f"no MUFU, no TMEM: {pl}")
    run(1 | pack, 0, 208, f"MUFU+FFMA, no TMEM: {pl}")
    run(2 | pack, 0, 208, f"TMEM ld/st only: {pl}")
    run(3 | pack, 0, 208, f"full epilogue: {pl}")
    nm = int(reps * 2200 / 104)
    run(3 | 4 | pack, nm, 208, f"full epilogue + MMAs N=208: {pl}")
run(3 | 16, 0, 208, "full epilogue, prefetched ld: F2FP pack")
run(3 | 16 | 4, int(reps * 2200 / 104), 208, "full epilogue, prefetched ld + MMAs: F2FP pack")
run(3 | 16 | 32, 0, 208, "full epilogue, prefetched ld: int half-up pack")
run(3 | 16 | 32 | 4, int(reps * 2200 / 104), 208, "full epilogue, prefetched ld + MMAs: int half-up pack")
run(4, int(reps * 2200 / 104), 208, "MMAs N=208 alone (epilogue loop: F2FP pack only)")
