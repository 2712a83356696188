#!/usr/bin/env python
"""bench.py -- candidate action-sequences/sec of the PETS HalfCheetah CEM planning loop (BASELINE.json).

One "step" = one full CEM plan, i.e. one ``TrajectoryOptimizerAgent.act()`` worth of work: 5 iterations x
population 500 candidate sequences, each rolled out for H = 30 steps with 20 particles through the ensemble
(7 members, 5 elites, 4 x 200 SiLU).  ``value`` times it device-resident (inputs in HBM, CUDA events);
``e2e`` times the public API call ``agent.act(obs)`` with a host observation in and the host plan out.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (under torchrun): weak scaling -- every rank plans over its own 500-sequence shard of a 500 x N
population, ONE all-gather of local top-k records per CEM iteration (mbrl_lib_b200.dist); ``e2e`` is then
``ShardedCEMOptimizer.optimize`` per call with the host observation in and the host plan out, max over ranks.
``--impl reference``: the oracle port of the reference's PyTorch path timed on the host CPUs (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mbrl_lib_b200 import synthetic as syn  # noqa: E402

WORKLOAD = "halfcheetah"  # BASELINE.json configs[1]
CEM_ITERS, ELITE_RATIO, ALPHA = 5, 0.1, 0.1
FLOP_PER_SEQ = 157.68e6  # 2 * sum K*N (true dims) * P * H = 262 800 * 20 * 30
METRIC = "candidate action-sequences/sec, PETS HalfCheetah CEM"
# the same workload string on both arms (ours and --impl reference)
WORKLOAD_DESC = ("PETS gym___HalfCheetah-v4 dims (obs 17, act 6): ensemble 7 (5 elites) x 4x200 SiLU, CEM pop 500 x 5 iterations "
                 "per plan, horizon 30, 20 particles, TS1")
NCU_DRAM_BYTES_PER_LAUNCH = 1936896 + 188416  # measured once per change with ncu (see profiles/), not at bench time


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_problem(device):
    import mbrl_lib_b200 as bp
    from mbrl_lib_b200 import functions

    spec = syn.CASES[WORKLOAD]
    arrays = syn.make_model_arrays(spec)
    model = bp.model_from_arrays(spec, arrays, device)

    class _Box:
        def __init__(self, lo, hi, n):
            self.low, self.high, self.shape = np.full(n, lo, np.float32), np.full(n, hi, np.float32), (n,)

    class _Env:
        observation_space = _Box(-np.inf, np.inf, spec.obs_dim)
        action_space = _Box(spec.action_lb, spec.action_ub, spec.act_dim)

    env = bp.ModelEnv(_Env(), model, functions.TERM_FNS[spec.term_fn], functions.REWARD_FNS[spec.reward_fn],
                      generator=torch.Generator(device=device).manual_seed(0), precision="auto", ts1="tile_shuffle")
    return spec, arrays, env


def l2_flush_buffer(device):
    return torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)


def run_ours(args):
    import torch.distributed as dist
    import mbrl_lib_b200 as bp
    from mbrl_lib_b200.planning import _FusedObjective

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    spec, arrays, env = build_problem(device)
    inp = syn.make_rollout_inputs(spec, with_noise=False)
    H, A, N, P = spec.horizon, spec.act_dim, spec.population, spec.particles
    lb = np.full((H, A), spec.action_lb).tolist()
    ub = np.full((H, A), spec.action_ub).tolist()
    obs0 = inp["obs0"]
    x0 = torch.zeros(H, A, device=device)
    flush = l2_flush_buffer(device)
    seqs_per_step = CEM_ITERS * N * world

    if world == 1:
        opt = bp.CEMOptimizer(CEM_ITERS, ELITE_RATIO, N, lb, ub, ALPHA, device, return_mean_elites=True)
        obj = _FusedObjective(env, obs0, P)

        def step():
            return opt.optimize(obj, x0=x0)
        launches_per_step = 1 + CEM_ITERS * 4  # init + (sample, rollout, particle mean, select/refit) per iteration
    else:
        from mbrl_lib_b200.dist import ShardedCEMOptimizer

        opt = ShardedCEMOptimizer(CEM_ITERS, ELITE_RATIO, N * world, lb, ub, ALPHA, device, return_mean_elites=True)

        obj = _FusedObjective(env, obs0, P)

        def step():
            return opt.optimize(obj, x0=x0)
        launches_per_step = CEM_ITERS * 5  # sample, rollout, particle mean, local top-k, refit-from-records

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: per-step CUDA events, L2 flushed (untimed) between steps ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in evs:
        flush.fill_(1)
        s.record()
        step()
        e.record()
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in evs]
    total_ms = torch.tensor([sum(step_ms)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())

    # ---- dominant kernel (the rollout) alone, for the roofline ----
    pop = torch.from_numpy(syn.make_rollout_inputs(spec, with_noise=False)["actions"]).to(device)
    for _ in range(3):
        env.evaluate_action_sequences(pop, obs0, P)
    torch.cuda.synchronize()
    nk = max(args.steps, 10)
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    for _ in range(nk):
        flush.fill_(1)
        ks.record()
        env.evaluate_action_sequences(pop, obs0, P)
        ke.record()
        torch.cuda.synchronize()
        kern_ms.append(ks.elapsed_time(ke))
    kern_ms_avg = sum(kern_ms) / len(kern_ms)

    # ---- end to end through the public API: agent.act(obs): H2D obs (pinned) ... D2H plan ----
    e2e = None
    if world == 1:
        cfg = {"_target_": "mbrl_lib_b200.TrajectoryOptimizerAgent", "planning_horizon": H, "replan_freq": 1, "verbose": False,
               "optimizer_cfg": {"_target_": "mbrl_lib_b200.CEMOptimizer", "num_iterations": CEM_ITERS,
                                 "elite_ratio": ELITE_RATIO, "population_size": N, "alpha": ALPHA, "device": device,
                                 "return_mean_elites": True}}
        agent = bp.create_trajectory_optim_agent_for_model(env, cfg, num_particles=P)
        for _ in range(3):
            agent.act(obs0)
        torch.cuda.synchronize()
        dt = 0.0
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()  # the L2 flush is not part of the measured call
            t0 = time.perf_counter()
            a = agent.act(obs0)  # host numpy in (H2D from pinned memory), host numpy out (synchronises on the D2H of the plan)
            dt += time.perf_counter() - t0
        assert a.shape == (A,)
        e2e = {"value": CEM_ITERS * N * args.steps / dt, "unit": "sequences/s", "h2d_bytes_per_step": spec.obs_dim * 4,
               "d2h_bytes_per_step": H * A * 4, "ms_per_step": dt / args.steps * 1e3,
               "note": "agent.act(obs) per call, host wall clock; L2 flushed (untimed) before every call"}
    if world > 1:
        # sharded plan through the public API: host observation in (pinned H2D inside optimize), host plan out
        pin = torch.empty(H, A, dtype=torch.float32).pin_memory()
        for _ in range(2):
            pin.copy_(opt.optimize(obj, x0=x0), non_blocking=True)
            torch.cuda.synchronize()
        dt = 0.0
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            t0 = time.perf_counter()
            pin.copy_(opt.optimize(obj, x0=x0), non_blocking=True)
            torch.cuda.synchronize()
            dt += time.perf_counter() - t0
        dt_t = torch.tensor([dt], device=device, dtype=torch.float64)
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t.item())
        e2e = {"value": seqs_per_step * args.steps / dt, "unit": "sequences/s", "h2d_bytes_per_step": spec.obs_dim * 4,
               "d2h_bytes_per_step": H * A * 4, "ms_per_step": dt / args.steps * 1e3,
               "note": "ShardedCEMOptimizer.optimize per call on every rank (max over ranks), host wall clock; L2 flushed before"}
    # ---- population scan of the rollout alone (BASELINE.json config 5 shape, one GPU): fills all SMs ----
    scan = []
    if world == 1 and not args.no_scan:
        peak_tf, _ = measured_peak_tflops()
        for scale in (2, 8, 32, 128):
            big = pop.repeat(scale, 1, 1)
            for _ in range(2):
                env.evaluate_action_sequences(big, obs0, P)
            torch.cuda.synchronize()
            reps = 5
            ks.record()
            for _ in range(reps):
                env.evaluate_action_sequences(big, obs0, P)
            ke.record()
            torch.cuda.synchronize()
            ms = ks.elapsed_time(ke) / reps
            sps = big.shape[0] / (ms * 1e-3)
            scan.append({"population": int(big.shape[0]), "ms": ms, "sequences_per_s": sps,
                         "tensor_roofline_frac": sps * FLOP_PER_SEQ / 1e12 / peak_tf})
            del big
    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(spec, arrays, reps=3)

    if rank == 0:
        peak, peak_src = measured_peak_tflops()
        achieved = N * FLOP_PER_SEQ / (kern_ms_avg * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": seqs_per_step * args.steps / (total_ms * 1e-3), "unit": "sequences/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if env.precision == "bf16_tc" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC, "step": "one 5-iteration CEM plan (2 500 sequences per GPU)",
                       "ts1": "in-kernel tile shuffle", "population_per_gpu": N, "sequences_per_step": seqs_per_step, "parallelism": f"population-sharded x{world}",
                       "l2": "flushed between timed steps (256 MB write, untimed); weights (1.3 MB) are re-fetched every step",
                       "precision": env.precision},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one "
                         "launch (profiles/r1_rollout_tc_final_ncu_full_summary.csv): 1.94 MB + 0.19 MB; algorithmic: 1.3 MB bf16 weights + "
                         "0.36 MB actions + 2 KB returns", "kernel": "rollout_tc_kernel (one CEM iteration: 500 sequences x 20 particles x 30 steps)",
                         "kernel_ms": kern_ms_avg, "peak_source": peak_src,
                         "algorithmic_flop_per_launch": N * FLOP_PER_SEQ},
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if scan:
            line["population_scan_rollout_only"] = scan
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def oracle_eval_once(model, spec, actions, obs0, gen):
    """One evaluate_action_sequences of the oracle port with the RNG work the reference does per step
    (randperm of B rows + N(0,1) draws of [B, out]) inside the timed region."""
    H, B = spec.horizon, spec.population * spec.particles
    perms = torch.stack([torch.randperm(B, generator=gen) for _ in range(H)])
    eps = torch.randn(H, B, spec.out_size, generator=gen)
    return model.evaluate_action_sequences(actions, obs0, spec.particles, perms, eps)


def pick_cpu_threads(model, spec, actions, obs0):
    """torch's default (one thread per hardware thread) oversubscribes these small GEMMs badly on a many-core host;
    time a 50-sequence slice at a few thread counts and keep the fastest."""
    cands = sorted({c for c in (8, 16, 32, 64, os.cpu_count() or 1) if c <= (os.cpu_count() or 1)})
    small = actions[:50]
    best, best_t = cands[0], float("inf")
    gen = torch.Generator().manual_seed(1)
    sub = type(spec)(**{**spec.__dict__, "population": 50})
    for c in cands:
        torch.set_num_threads(c)
        oracle_eval_once(model, sub, small, obs0, gen)
        t0 = time.perf_counter()
        oracle_eval_once(model, sub, small, obs0, gen)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline(spec, arrays, reps=3):
    from oracle import pets_oracle as po

    model = po.OracleModel(spec, arrays)
    inp = syn.make_rollout_inputs(spec, with_noise=False)
    actions = torch.from_numpy(inp["actions"])
    threads = pick_cpu_threads(model, spec, actions, inp["obs0"])
    gen = torch.Generator().manual_seed(0)
    oracle_eval_once(model, spec, actions, inp["obs0"], gen)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle_eval_once(model, spec, actions, inp["obs0"], gen)
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": spec.population / med, "unit": "sequences/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x evaluate_action_sequences of 500 sequences (one CEM iteration), oracle port of the "
                      f"reference's fp32 PyTorch path, best of thread counts up to {os.cpu_count()}, median {med:.3f} s"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port; the Python reference itself
    cannot travel to the GPU box) on all host threads.  One step = one CEM iteration's evaluation (500 sequences)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pets_oracle as po

    spec = syn.CASES[WORKLOAD]
    arrays = syn.make_model_arrays(spec)
    model = po.OracleModel(spec, arrays)
    inp = syn.make_rollout_inputs(spec, with_noise=False)
    actions = torch.from_numpy(inp["actions"])
    pick_cpu_threads(model, spec, actions, inp["obs0"])
    gen = torch.Generator().manual_seed(0)
    steps = min(args.steps, 40)
    for _ in range(min(max(args.warmup, 1), 3)):
        oracle_eval_once(model, spec, actions, inp["obs0"], gen)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_eval_once(model, spec, actions, inp["obs0"], gen)
    dt = time.perf_counter() - t0
    val = spec.population * steps / dt
    cores = torch.get_num_threads()
    sample = (f"{steps} x evaluate_action_sequences of 500 sequences (one CEM iteration each), {cores} threads "
              f"(fastest of the thread counts tried on {os.cpu_count()} hardware threads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "sequences/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": min(max(args.warmup, 1), 3), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC, "step": "bounded sample: one CEM iteration's evaluation (500 sequences)",
                   "ts1": "torch.randperm per step (the reference's rule)", "device": "cpu"},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-scan", action="store_true", help="skip the population scan of the rollout kernel")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
