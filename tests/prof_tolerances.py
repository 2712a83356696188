"""Diagnostics: the actual size of the deviations behind the two loose bars of tests/test_gpu_parity.py (discrete rewards on the
tensor-core kernel; bf16 fused-plan solution), against the fp32 golden AND against the bf16-operand oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import test_gpu_parity as tp
from mbrl_lib_b200 import synthetic as syn

gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
for name in ["cartpole", "hopper_tsinf", "walker_ant", "ant_learned_fn", "relu_expectation"]:
    spec, arrays, env = tp.make_env(name, "bf16_tc")
    inp = syn.make_rollout_inputs(spec)
    got = tp.gpu_returns(env, spec, inp)
    gold = np.load(os.path.join(gd, f"rollout_{name}.npz"))["returns"]
    orc = tp.oracle_returns(spec, arrays, inp, bf16=True)
    for tag, ref in (("fp32 golden", gold), ("bf16 oracle", orc)):
        diff = np.abs(got - ref)
        frac = float((diff > 1e-2 * np.maximum(1.0, np.abs(ref))).mean())
        print(f"{name:18s} vs {tag}: frac(|diff| > 1e-2 scale) = {frac:.4f}, mean diff = {diff.mean():.4e}, "
              f"|mean got - mean ref| = {abs(got.mean() - ref.mean()):.4e}, n = {diff.size}")

# fused CEM plan, bf16: solution against the fp32 golden
import mbrl_lib_b200 as bp
from mbrl_lib_b200.planning import _FusedObjective
g = np.load(os.path.join(gd, "cem_model.npz"))
for precision in ("f32", "bf16_tc"):
    spec, arrays, env = tp.make_env("halfcheetah_small", precision)
    inp = syn.make_rollout_inputs(spec)
    iters = int(g["iters"])
    nz = syn.make_cem_noise(spec, iters)
    H, A = spec.horizon, spec.act_dim
    lb = np.full((H, A), spec.action_lb).tolist(); ub = np.full((H, A), spec.action_ub).tolist()
    opt = bp.CEMOptimizer(iters, 0.1, spec.population, lb, ub, 0.1, tp.DEV, return_mean_elites=True)
    opt.record_values = True
    sol = opt.optimize(_FusedObjective(env, inp["obs0"], spec.particles), x0=torch.zeros(H, A, device=tp.DEV),
                       _noise=torch.from_numpy(nz["z"]).to(tp.DEV),
                       _model_noise=(torch.from_numpy(nz["perms"]).to(tp.DEV), torch.from_numpy(nz["eps"]).to(tp.DEV)))
    torch.cuda.synchronize()
    vals = opt.last_values.cpu().numpy()
    scale = max(1.0, np.abs(g["values"]).max())
    print(f"fused plan {precision}: max |solution - golden| = {np.abs(sol.cpu().numpy() - g['solution']).max():.4e}, "
          f"values iteration 0: {np.abs(vals[0] - g['values'][0]).max() / scale:.3e} of scale, all iterations: {np.abs(vals - g['values']).max() / scale:.3e}")
