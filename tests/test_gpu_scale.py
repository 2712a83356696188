"""Full-size / size-independent properties on the GPU (BASELINE.json configs 2, 4, 5) where the oracle is too slow to
be the checker: determinism, top-k invariants of the radix-select path, conservation checks of the MBPO step."""
import numpy as np
import pytest
import torch

from mbrl_lib_b200 import synthetic as syn
from test_gpu_parity import DEV, make_env

pytestmark = pytest.mark.gpu


def test_large_population_refit_matches_torch():
    """config 5 scale: N = 64 000, k = 6 400 goes through the radix-select path; torch.topk / mean / var on the device
    is the checker (library code used as a test oracle only)."""
    import mbrl_lib_b200 as bp
    from mbrl_lib_b200 import _lib

    lib = _lib.load()
    N, dims, k, alpha = 64000, 36, 6400, 0.1
    g = torch.Generator(device=DEV).manual_seed(3)
    pop = torch.randn(N, dims, device=DEV, generator=g)
    vals = torch.randn(N, device=DEV, generator=g)
    vals[::977] = float("nan")
    vals[5::1000] = 0.25  # ties
    mu = torch.zeros(dims, device=DEV)
    disp = torch.ones(dims, device=DEV)
    best_v = torch.full((1,), float("-inf"), device=DEV)
    best_s = torch.zeros(dims, device=DEV)
    idx = torch.empty(k, dtype=torch.int32, device=DEV)
    nbytes = lib.b200pets_cem_update_workspace_bytes(N, dims, k)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    v_in = vals.clone()
    _lib.check(lib.b200pets_cem_update(N, dims, k, alpha, 1, 0, _lib.ptr(pop), _lib.ptr(v_in), _lib.ptr(mu), _lib.ptr(disp),
                                       _lib.ptr(best_v), _lib.ptr(best_s), _lib.ptr(idx), None, _lib.ptr(ws), nbytes,
                                       _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref_v = vals.clone()
    ref_v[ref_v.isnan()] = -1e-10
    assert torch.equal(v_in, ref_v)  # NaN rule applied in place
    top_v, _ = ref_v.topk(k)
    sel = idx.long()
    assert sel.unique().numel() == k and bool((sel[1:] > sel[:-1]).all())  # k distinct indices, ascending
    thr = top_v[-1]
    assert bool((ref_v[sel] >= thr).all())
    assert torch.allclose(ref_v[sel].sort(descending=True).values, top_v)  # same multiset of values as torch.topk
    elite = pop[sel]
    torch.testing.assert_close(mu, (1 - alpha) * elite.mean(0), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(disp, alpha * torch.ones(dims, device=DEV) + (1 - alpha) * elite.var(0), rtol=1e-4, atol=1e-5)
    assert float(best_v) == float(ref_v.max())
    torch.testing.assert_close(best_s, pop[int(ref_v.argmax())])


@pytest.mark.parametrize("precision", ["bf16_tc", "f32"])
def test_rollout_is_deterministic_at_scale(precision):
    spec, arrays, env = make_env("halfcheetah", precision, ts1="tile_shuffle")
    N = 4000 if precision == "bf16_tc" else 1000
    g = np.random.default_rng(0)
    acts = torch.from_numpy(g.uniform(-1, 1, (N, spec.horizon, spec.act_dim)).astype(np.float32)).to(DEV)
    obs0 = syn.make_rollout_inputs(spec, with_noise=False)["obs0"]
    env._offset = 100
    r1 = env.evaluate_action_sequences(acts, obs0, spec.particles)
    env._offset = 100
    r2 = env.evaluate_action_sequences(acts, obs0, spec.particles)
    assert torch.equal(r1, r2) and bool(torch.isfinite(r1).all())  # same Philox (seed, offset) => bit-identical
    env._offset = 101
    r3 = env.evaluate_action_sequences(acts, obs0, spec.particles)
    assert not torch.equal(r1, r3)


@pytest.mark.parametrize("precision", ["bf16_tc", "f32"])
def test_mbpo_step_full_size(precision):
    """config 4: 100 000 start states x 1 step (mbpo.py:31-63).  Mean prediction (sample=False) must equal the model
    applied row by row: check a random subset against the oracle, and conservation properties on all rows."""
    from oracle import pets_oracle as po

    spec, arrays, env = make_env("mbpo_halfcheetah", precision, ts1="tile_shuffle")
    B = 100000
    inp = syn.make_step_inputs(spec, B)
    state = env.reset(inp["obs"], return_as_np=False)
    act = torch.from_numpy(inp["act"]).to(DEV)
    perm = torch.from_numpy(inp["perm"]).to(DEV)
    nobs, rew, done, _ = env.step(act, state, sample=False, _perm=perm)
    assert nobs.shape == (B, spec.obs_dim) and rew.shape == (B, 1) and done.shape == (B, 1)
    assert bool(torch.isfinite(nobs).all()) and not bool(done.any())  # no_termination
    # a slice of member 0's rows, checked against the oracle's member MLP
    M = len(spec.elites)
    rows = perm[: B // M][:256].cpu()
    m = po.OracleModel(spec, arrays)
    m.emulate_bf16 = precision == "bf16_tc"
    x = m.model_input(torch.from_numpy(inp["obs"])[rows], torch.from_numpy(inp["act"])[rows])
    mean, _ = m.mlp(x.unsqueeze(0).expand(M, -1, -1).contiguous())
    pred = mean[0]
    ref_nobs = pred[:, :-1] + torch.from_numpy(inp["obs"])[rows]
    tol = 2e-4 if precision == "f32" else 5e-3
    scale = max(1.0, float(ref_nobs.abs().max()))
    assert float((nobs[rows.to(DEV)].cpu() - ref_nobs).abs().max()) <= tol * scale
    assert float((rew[rows.to(DEV)].cpu()[:, 0] - pred[:, -1]).abs().max()) <= tol * scale


def test_icem_over_model_runs_and_improves():
    """config 3 shape family: iCEM (coloured noise, decaying population, kept elites) driving the model rollout."""
    import mbrl_lib_b200 as bp

    spec, arrays, env = make_env("humanoid_trunc", "auto", ts1="tile_shuffle")
    H, A = spec.horizon, spec.act_dim
    lb, ub = np.full((H, A), spec.action_lb).tolist(), np.full((H, A), spec.action_ub).tolist()
    opt = bp.ICEMOptimizer(4, 0.1, 350, 1.3, 2.0, lb, ub, 0.3, 0.1, DEV, return_mean_elites=False, population_size_module=5)
    obs0 = syn.make_rollout_inputs(spec, with_noise=False)["obs0"]
    best = []

    def obj(pop):
        return env.evaluate_action_sequences(pop, obs0, spec.particles)

    sol = opt.optimize(obj, x0=torch.zeros(H, A, device=DEV), callback=lambda p, v, i: best.append(float(v.max())))
    assert sol.shape == (H, A) and bool(torch.isfinite(sol).all())
    assert float(sol.min()) >= spec.action_lb - 1e-6 and float(sol.max()) <= spec.action_ub + 1e-6
    assert opt.elite.shape == (opt.elite_num, H, A)
    assert max(best[1:]) >= best[0] - 0.05  # later generations are at least as good as the first (noisy objective)
    sol2 = opt.optimize(obj, x0=torch.zeros(H, A, device=DEV))  # second call: shifted elites of the previous plan
    assert bool(torch.isfinite(sol2).all())


def _line_world(device):
    """A hand-built deterministic 'ensemble' that is exactly a 1-D point mass (ReLU pairs encode identity / abs):
    next_pos = pos + 0.1 * a, learned reward = -|next_pos|.  Plays the role of the reference's MockLineEnv
    integration test (tests/algorithms/test_algorithms.py:44-68) without needing model training."""
    import mbrl_lib_b200 as bp

    mlp = bp.GaussianMLP(2, 2, device, num_layers=1, ensemble_size=2, hid_size=32, deterministic=True,
                         propagation_method="random_model", activation="relu")
    w0 = torch.zeros(2, 2, 32)
    w0[:, 1, 0], w0[:, 1, 1] = 1.0, -1.0                      # relu(a), relu(-a)
    w0[:, 0, 2], w0[:, 1, 2] = 1.0, 0.1                       # relu(pos + 0.1 a)
    w0[:, 0, 3], w0[:, 1, 3] = -1.0, -0.1                     # relu(-(pos + 0.1 a))
    w1 = torch.zeros(2, 32, 2)
    w1[:, 0, 0], w1[:, 1, 0] = 0.1, -0.1                      # delta = 0.1 a
    w1[:, 2, 1], w1[:, 3, 1] = -1.0, -1.0                     # reward = -|pos + 0.1 a|
    with torch.no_grad():
        mlp.hidden_layers[0][0].weight.copy_(w0)
        mlp.hidden_layers[0][0].bias.zero_()
        mlp.mean_and_logvar.weight.copy_(w1)
        mlp.mean_and_logvar.bias.zero_()
    return bp.OneDTransitionRewardModel(mlp, target_is_delta=True, normalize=False, learned_rewards=True)


@pytest.mark.parametrize("precision", ["f32", "bf16_tc"])
def test_closed_loop_mpc_reaches_goal(precision):
    """Behavioural check of the whole agent loop on the GPU (act -> fused CEM plan -> warm-start shift -> act ...)."""
    import mbrl_lib_b200 as bp
    from mbrl_lib_b200 import functions
    from test_gpu_parity import _Env

    class _Spec:
        obs_dim, act_dim, action_lb, action_ub = 1, 1, -1.0, 1.0

    model = _line_world(DEV)
    env = bp.ModelEnv(_Env(_Spec), model, functions.no_termination, None, generator=torch.Generator(device=DEV),
                      precision=precision, ts1="tile_shuffle")
    cfg = {"_target_": "mbrl.planning.TrajectoryOptimizerAgent", "planning_horizon": 5, "replan_freq": 1,
           "optimizer_cfg": {"_target_": "mbrl.planning.CEMOptimizer", "num_iterations": 4, "elite_ratio": 0.1,
                             "population_size": 256, "alpha": 0.1, "device": DEV, "return_mean_elites": True}}
    agent = bp.create_trajectory_optim_agent_for_model(env, cfg, num_particles=2)
    pos, total = 1.0, 0.0
    for _ in range(18):
        a = float(np.clip(agent.act(np.array([pos])), -1, 1)[0])
        pos = pos + 0.1 * a  # the true environment is the same point mass
        total += -abs(pos)
    assert abs(pos) < 0.12, pos            # started at 1.0, can move 0.1 per step
    assert total > -7.5, total             # an agent that never moves collects -18


@pytest.mark.gpu
def test_rollout_model_env_matches_oracle_open_loop():
    """util/common.py:416-454 on the CUDA step kernel vs. the oracle stepping the same plan (TSinf, sample=False)."""
    import mbrl_lib_b200 as bp
    from oracle import pets_oracle as po

    spec, arrays, env = make_env("hopper_tsinf", "f32")
    S, L = 3 * spec.num_models, 6
    rng = np.random.default_rng(5)
    obs0 = rng.standard_normal(spec.obs_dim).astype(np.float32)
    plan = rng.uniform(spec.action_lb, spec.action_ub, size=(L, spec.act_dim)).astype(np.float32)
    torch.manual_seed(11)
    obs, rew, got_plan = bp.rollout_model_env(env, obs0, plan, None, num_samples=S)
    assert obs.shape == (L + 1, S, spec.obs_dim) and rew.shape == (L, S, 1) and got_plan is plan
    torch.manual_seed(11)
    perm = torch.randperm(S, device="cuda").cpu()  # the draw ModelEnv.reset made
    m = po.OracleModel(spec, arrays)
    o = torch.from_numpy(np.tile(obs0, (S, 1)))
    for t in range(L):
        o, r, _ = m.step(o, torch.from_numpy(np.tile(plan[t], (S, 1))), perm, None, sample=False)
        scale = max(1.0, float(o.abs().max()))
        assert np.abs(obs[t + 1] - o.numpy()).max() <= 5e-4 * scale
        assert np.abs(rew[t] - r.numpy().reshape(S, 1)).max() <= 5e-4 * max(1.0, float(r.abs().max()))
