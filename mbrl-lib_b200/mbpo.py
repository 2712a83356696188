"""MBPO model rollouts kept on the device: ``rollout_model_and_populate_sac_buffer`` with the reference's signature
(mbrl/algorithms/mbpo.py:31-63).

The reference moves every step's ``(next_obs, rewards, dones)`` to the host, masks them with numpy and calls
``sac_buffer.add_batch`` per step.  Here the observation batch, the predictions of all ``rollout_horizon`` steps and the
``accum_dones`` mask stay in HBM; one ordered compaction (``b200pets_mbpo_compact``) packs the alive transitions of all
steps, ONE device->host copy brings them back, and ``add_batch`` is then called once per step on host slices -- the
replay buffer ends up with exactly the rows, in exactly the order, the reference would have stored.

The policy: ``agent.act(obs, sample=..., batched=True)`` is numpy-in / numpy-out in the reference
(mbrl/planning/sac_wrapper.py:27-46).  To keep the loop on the device the actor is taken as an opaque torch callable:
``agent.act_torch(obs_tensor, sample)`` when the agent offers it, else ``agent.sac_agent.policy.sample`` (the
pytorch_sac policy the reference's SACAgent wraps), else the numpy ``agent.act`` with a per-step round trip of the
observations / actions only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _policy(agent, device):
    if hasattr(agent, "act_torch"):
        return lambda obs, sample: agent.act_torch(obs, sample)
    sac = getattr(agent, "sac_agent", None)
    if sac is not None and hasattr(sac, "policy") and hasattr(sac.policy, "sample"):
        def f(obs, sample):  # pytorch_sac select_action: (action, log_prob, mean) = policy.sample(state)
            with torch.no_grad():
                action, _, mean = sac.policy.sample(obs.to(getattr(sac, "device", device)))
            return (action if sample else mean).to(device)
        return f

    def g(obs, sample):  # opaque numpy agent: only the observations / actions cross PCIe
        a = agent.act(obs.cpu().numpy(), sample=sample, batched=True)
        return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(device)
    return g


def rollout_on_device(model_env, initial_obs: np.ndarray, agent, sac_samples_action: bool, rollout_horizon: int, *,
                      _noise=None, _staging=None):
    """The device part of the rollout: returns host arrays ``(obs, act, next_obs, reward, done, counts)`` holding the
    alive transitions of all steps packed in (step, row) order; ``counts[i]`` rows belong to step i.
    (``_noise``: per-step ``(perm, eps)`` device tensors injected into ModelEnv.step; ``_staging``: a dict that receives
    the un-compacted device buffers -- both for parity tests.)"""
    lib = _lib.load()
    dev = model_env.device
    k = int(rollout_horizon)
    state = model_env.reset(np.asarray(initial_obs), return_as_np=False)
    obs0 = state["obs"]
    B, D = obs0.shape
    A = int(model_env.action_space.shape[0])
    policy = _policy(agent, dev)
    act = torch.empty(k, B, A, device=dev)
    nxt = torch.empty(k, B, D, device=dev)
    rew = torch.empty(k, B, device=dev)
    done = torch.empty(k, B, dtype=torch.uint8, device=dev)
    alive = torch.empty(k, B, dtype=torch.uint8, device=dev)
    accum = torch.zeros(B, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        for i in range(k):
            a = policy(state["obs"], sac_samples_action)
            act[i].copy_(a.reshape(B, A))
            pn, en = _noise[i] if _noise is not None else (None, None)
            model_env.step(act[i], state, sample=True, _out=(nxt[i], rew[i], done[i]), _perm=pn, _eps=en)
            _lib.check(lib.b200pets_mbpo_mask(B, _lib.ptr(done[i]), _lib.ptr(accum), _lib.ptr(alive[i]), _lib.stream_ptr()),
                       "mbpo_mask")
            state = dict(state)
            state["obs"] = nxt[i]
        o_out = torch.empty(k * B, D, device=dev)
        a_out = torch.empty(k * B, A, device=dev)
        n_out = torch.empty(k * B, D, device=dev)
        r_out = torch.empty(k * B, device=dev)
        d_out = torch.empty(k * B, dtype=torch.uint8, device=dev)
        counts = torch.empty(k + 1, dtype=torch.int64, device=dev)
        need = lib.b200pets_mbpo_compact_workspace_bytes(k, B)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _lib.check(lib.b200pets_mbpo_compact(k, B, D, A, _lib.ptr(obs0), _lib.ptr(act), _lib.ptr(nxt), _lib.ptr(rew),
                                             _lib.ptr(done), _lib.ptr(alive), _lib.ptr(o_out), _lib.ptr(a_out), _lib.ptr(n_out),
                                             _lib.ptr(r_out), _lib.ptr(d_out), _lib.ptr(counts), _lib.ptr(ws), need,
                                             _lib.stream_ptr()), "mbpo_compact")
    if _staging is not None:
        _staging.update(obs0=obs0, act=act, next_obs=nxt, reward=rew, done=done, alive=alive)
    counts_h = counts.cpu().numpy()  # synchronises: the number of rows to bring back
    total = int(counts_h[k])
    host = [t[:total].cpu().numpy() for t in (o_out, a_out, n_out, r_out, d_out)]
    return (*host, counts_h[:k])


def rollout_model_and_populate_sac_buffer(model_env, replay_buffer, agent, sac_buffer, sac_samples_action: bool,
                                          rollout_horizon: int, batch_size: int):
    """Drop-in for mbrl/algorithms/mbpo.py:31-63."""
    batch = replay_buffer.sample(batch_size)
    initial_obs, *_ = batch.astuple()
    obs, act, nxt, rew, done, counts = rollout_on_device(model_env, initial_obs, agent, sac_samples_action, rollout_horizon)
    lo = 0
    for n in counts:  # one add_batch per step, as the reference issues them (replay_buffer.py:553-599 wraps per call)
        hi = lo + int(n)
        sac_buffer.add_batch(obs[lo:hi], act[lo:hi], nxt[lo:hi], rew[lo:hi], done[lo:hi].astype(bool),
                             np.zeros(hi - lo, dtype=bool))
        lo = hi
