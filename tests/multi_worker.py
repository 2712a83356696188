"""Worker of tests/test_gpu_multi.py (launched under torchrun, one rank per GPU, NCCL): ShardedCEMOptimizer over the ranks
must produce, on every rank, the plan CEMOptimizer produces on ONE GPU over the union population for the same seed
(population noise / model noise / member draws are keyed by global indices; records are gathered in global index order)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    import mbrl_lib_b200 as bp
    from mbrl_lib_b200 import functions, synthetic as syn
    from mbrl_lib_b200.dist import ShardedCEMOptimizer
    from mbrl_lib_b200.planning import _FusedObjective
    import test_gpu_parity as tp

    spec = syn.CASES["halfcheetah"]
    arrays = syn.make_model_arrays(spec)
    H, A, N, P = spec.horizon, spec.act_dim, spec.population, spec.particles
    lb, ub = np.full((H, A), spec.action_lb).tolist(), np.full((H, A), spec.action_ub).tolist()
    inp = syn.make_rollout_inputs(spec, with_noise=False)
    out = {}
    for precision, peer_mode in (("bf16_tc", "1"), ("f32", "1"), ("bf16_tc", "0")):
        # peer_mode 1: records exchanged over NVLink peer memory by the top-k kernel itself; 0: NCCL all-gather
        os.environ["B200PETS_PEER_EXCHANGE"] = peer_mode

        def make_env():
            model = bp.model_from_arrays(spec, arrays, dev)
            return bp.ModelEnv(tp._Env(spec), model, functions.TERM_FNS[spec.term_fn], functions.REWARD_FNS[spec.reward_fn],
                               generator=torch.Generator(device=dev), precision=precision, ts1="tile_shuffle")

        iters = 3
        env = make_env()
        opt = ShardedCEMOptimizer(iters, 0.1, N, lb, ub, 0.1, dev, return_mean_elites=True)
        opt.record_values = True
        sol = opt.optimize(_FusedObjective(env, inp["obs0"], P), x0=torch.zeros(H, A, device=dev))
        torch.cuda.synchronize()
        used_peers = opt._peer is not None
        if peer_mode == "1" and not used_peers:
            raise RuntimeError("peer exchange was not set up on a single-node NCCL group")
        # every rank holds the same plan without a broadcast
        gathered = [torch.empty_like(sol) for _ in range(world)]
        dist.all_gather(gathered, sol)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        vals = [torch.empty_like(opt.last_values) for _ in range(world)]
        dist.all_gather(vals, opt.last_values.contiguous())
        ok_ref = True
        if rank == 0:  # the single-GPU plan over the union population, same seed / call counter
            env1 = make_env()
            ref_opt = bp.CEMOptimizer(iters, 0.1, N, lb, ub, 0.1, dev, return_mean_elites=True)
            ref_opt.record_values = True
            ref = ref_opt.optimize(_FusedObjective(env1, inp["obs0"], P), x0=torch.zeros(H, A, device=dev))
            torch.cuda.synchronize()
            union = torch.cat(vals, dim=1)
            ok_ref = bool(torch.equal(union, ref_opt.last_values)) and bool(torch.equal(ref, sol))
            out[f"{precision}_{'peer' if used_peers else 'nccl'}"] = {"ranks_agree": same, "equals_single_gpu_plan": ok_ref,
                              "max_abs_diff_plan": float((ref - sol).abs().max()),
                              "max_abs_diff_values": float((union - ref_opt.last_values).abs().max())}
        flag = torch.tensor([int(same and ok_ref)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            if rank == 0:
                print(json.dumps({"multi_gpu_check": "FAILED", "world": world, **out}), flush=True)
            dist.destroy_process_group()
            sys.exit(1)
    if rank == 0:
        print(json.dumps({"multi_gpu_check": "ok", "world": world, **out}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
