"""Profiling target: a few evaluate_action_sequences calls at the headline config (in-kernel RNG, tile shuffle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mbrl_lib_b200 import synthetic as syn

pop_scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spec, arrays, env = bench.build_problem("cuda:0")
inp = syn.make_rollout_inputs(spec, with_noise=False)
acts = torch.from_numpy(inp["actions"]).to("cuda:0").repeat(pop_scale, 1, 1)
for _ in range(6):
    env.evaluate_action_sequences(acts, inp["obs0"], spec.particles)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    env.evaluate_action_sequences(acts, inp["obs0"], spec.particles)
e.record(); torch.cuda.synchronize()
print(f"pop {acts.shape[0]}: {s.elapsed_time(e)/10:.4f} ms per evaluation, {acts.shape[0]/(s.elapsed_time(e)/10*1e-3):.3e} seq/s")
