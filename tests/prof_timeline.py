"""Diagnostics: per-phase clock64 stamps of one horizon step of CTA 0 (tensor-core rollout)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mbrl_lib_b200 import synthetic as syn, _lib

spec, arrays, env = bench.build_problem("cuda:0")
inp = syn.make_rollout_inputs(spec, with_noise=False)
acts = torch.from_numpy(inp["actions"]).to("cuda:0")
lib = _lib.load()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda:0")
for _ in range(3):
    env.evaluate_action_sequences(acts, inp["obs0"], spec.particles)
lib.b200pets_debug_timeline(_lib.ptr(buf))
env.evaluate_action_sequences(acts, inp["obs0"], spec.particles)
torch.cuda.synchronize()
lib.b200pets_debug_timeline(None)
b = buf.cpu().tolist()
t0 = b[0]
print("epilogue thread stamps (cycles since step start):", [x - t0 for x in b[:24] if x])
print("fine stamps (slot u: 40+4u loads done, 41+4u noise ready, 42+4u state written, 43+4u operand part begins; 48+u operand stored; 52 before wait::st, 53 after, 54 arrived):",
      {i: x - t0 for i, x in enumerate(b[40:56], start=40) if x})
print("column split CS-1 (two output groups): output accumulator ready, groups done, next input handed over:", [x - t0 for x in b[56:59] if x])
for l in range(5):
    s = b[64 + 4 * l: 68 + 4 * l]
    print(f"mma layer {l}: starts waiting {s[0]-t0}, weights + first activation half ready {s[1]-t0}, second half seen {s[2]-t0}, all MMAs issued {s[3]-t0}")

print("per epilogue warp (q, cs), hidden layer l: h1 arrived(l-1 -> this row shows l's own), reached acc0 wait, passed it, h0 arrived")
for w in range(16):
    row = b[400 + 16 * w: 416 + 16 * w]
    print(f"  warp q{w % 4} cs{w // 4}: " + "  ".join(f"l{l}: wait@{row[4*l+1]-t0} pass@{row[4*l+2]-t0} h0@{row[4*l+3]-t0} h1@{row[4*l]-t0}" for l in range(4)))

for name, base, kb in (("CTA 0", 128, 256), ("CTA 40", 192, 264)):
    st = [x for x in b[base:base + 60] if x]
    k0, g0, k1, g1 = b[kb:kb + 4]
    if not st or not k0:
        continue
    d = [st[i + 1] - st[i] for i in range(len(st) - 1)]
    print(f"coarse {name}: kernel entry -> first step start {st[0] - k0} cycles; step periods {d}")
    print(f"coarse {name}: last step start -> CTA done {k1 - st[-1]} cycles; CTA total {k1 - k0} cycles = {g1 - g0} ns "
          f"=> {1e3 * (k1 - k0) / max(g1 - g0, 1):.0f} MHz")
