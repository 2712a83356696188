"""Staging a ``OneDTransitionRewardModel(GaussianMLP)`` for the kernels (``b200pets_model_create``).

Reads the live ``nn.Parameter`` s / normaliser tensors of the user's model object (mbrl-lib's or the
containers in :mod:`models`), never copies them to the host, and keeps the packed device copy as a cache
keyed on (parameter storage, version counters, elite list, normaliser tensor identity) because
``ModelTrainer.train`` mutates weights in place and ``Normalizer.update_stats`` *replaces* its tensors
between ``act()`` calls (mbrl/models/model_trainer.py:153,288-296; mbrl/util/math.py:114-127).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib, functions


def _activation_of(module) -> tuple:
    name = type(module).__name__
    if name == "ReLU":
        return _lib.ACT["relu"], 0.0
    if name == "SiLU":
        return _lib.ACT["silu"], 0.0
    if name == "LeakyReLU":
        return _lib.ACT["leaky_relu"], float(module.negative_slope)
    raise NotImplementedError(f"activation {name} has no device implementation (known: ReLU, SiLU, LeakyReLU)")


class StagedModel:
    """Owns the C handle of one staged model and re-stages it when the source object changed."""

    def __init__(self, dynamics_model, reward_fn=None, termination_fn=None, stage: bool = True):
        """``stage=False`` only reads the source object (description, layer list, members, signature) and touches
        neither the library nor a device: that is how the duck-typing contract is checked against real
        ``mbrl.models`` objects on a CPU-only host (tests/test_reference_objects.py)."""
        self.lib = _lib.load() if stage else None
        self.src = dynamics_model
        mlp = getattr(dynamics_model, "model", None)
        if mlp is None or not hasattr(mlp, "hidden_layers") or not hasattr(mlp, "mean_and_logvar"):
            raise NotImplementedError(
                "the fused path covers OneDTransitionRewardModel(GaussianMLP); got "
                f"{type(dynamics_model).__name__}({type(mlp).__name__ if mlp is not None else None})")
        self.mlp = mlp
        dev = torch.device(mlp.mean_and_logvar.weight.device)
        if stage and dev.type != "cuda":
            raise RuntimeError(f"b200pets runs on a CUDA device; the model lives on {dev} (no CPU fallback)")
        self.device = dev
        self.reward_id = functions.resolve_reward(reward_fn)
        self.term_id = functions.resolve_term(termination_fn) if termination_fn is not None else _lib.TERM["no_termination"]
        self.handle: Optional[C.c_void_p] = None
        self._sig = None
        self._structure = None
        if stage:
            self.ensure_fresh()

    # ---- description -----------------------------------------------------------------------------------
    def _layers(self) -> List:
        return [seq[0] for seq in self.mlp.hidden_layers] + [self.mlp.mean_and_logvar]

    def members(self) -> List[int]:
        el = getattr(self.mlp, "elite_models", None)
        return list(el) if el is not None else list(range(int(self.mlp.num_members)))

    def _describe(self) -> _lib.ModelDesc:
        m, w = self.mlp, self.src
        layers = self._layers()
        act, slope = _activation_of(self.mlp.hidden_layers[0][1])
        d = _lib.ModelDesc()
        d.ensemble_size = int(layers[0].weight.shape[0])
        d.num_members = len(self.members())
        d.in_size = int(m.in_size)
        d.out_size = int(m.out_size)
        d.hid_size = int(layers[0].weight.shape[2])
        d.num_hidden = len(layers) - 1
        d.activation, d.leaky_slope = act, slope
        d.obs_process = functions.resolve_obs_process(getattr(w, "obs_process_fn", None))
        d.learned_rewards = int(bool(w.learned_rewards))
        d.target_is_delta = int(bool(w.target_is_delta))
        d.deterministic = int(bool(m.deterministic))
        d.obs_dim = d.out_size - d.learned_rewards
        d.act_dim = d.in_size - d.obs_dim - (1 if d.obs_process == _lib.PROC["cartpole"] else 0)
        d.reward_fn = self.reward_id
        d.term_fn = self.term_id
        norm = getattr(w, "input_normalizer", None)
        d.norm_mode = 0 if norm is None else (2 if norm.mean.dtype == torch.float64 else 1)
        if not d.learned_rewards and d.reward_fn == _lib.REWARD["learned"]:
            raise ValueError("reward_fn is None but the model does not learn rewards")
        return d

    def _signature(self):
        sig = []
        for layer in self._layers():
            for p in (layer.weight, layer.bias):
                sig.append((p.data_ptr(), p._version, tuple(p.shape)))
        if not self.mlp.deterministic:
            for p in (self.mlp.min_logvar, self.mlp.max_logvar):
                sig.append((p.data_ptr(), p._version))
        norm = getattr(self.src, "input_normalizer", None)
        if norm is not None:
            sig.append((id(norm.mean), norm.mean.data_ptr(), norm.mean._version, id(norm.std), norm.std._version))
        sig.append(tuple(self.members()))
        return tuple(sig)

    # ---- staging ---------------------------------------------------------------------------------------
    def ensure_fresh(self):
        sig = self._signature()
        if sig == self._sig:
            return
        desc = self._describe()
        structure = tuple(getattr(desc, f[0]) for f in desc._fields_) + tuple(getattr(self.src, "no_delta_list", []) or [])
        layers = self._layers()
        for layer in layers:
            for p in (layer.weight, layer.bias):
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != self.device:
                    raise ValueError("ensemble weights must be contiguous float32 tensors on one CUDA device")
        n = len(layers)
        W = (C.c_void_p * n)(*[layer.weight.data_ptr() for layer in layers])
        Bv = (C.c_void_p * n)(*[layer.bias.data_ptr() for layer in layers])
        members = self.members()
        mem = (C.c_int32 * len(members))(*members)
        norm = getattr(self.src, "input_normalizer", None)
        nm = ns = None
        if norm is not None:
            nm_np = np.ascontiguousarray(norm.mean.detach().double().cpu().numpy().reshape(-1))
            ns_np = np.ascontiguousarray(norm.std.detach().double().cpu().numpy().reshape(-1))
            nm = nm_np.ctypes.data_as(C.POINTER(C.c_double))
            ns = ns_np.ctypes.data_as(C.POINTER(C.c_double))
        mn = mx = None
        if not self.mlp.deterministic:
            mn_np = np.ascontiguousarray(self.mlp.min_logvar.detach().float().cpu().numpy().reshape(-1))
            mx_np = np.ascontiguousarray(self.mlp.max_logvar.detach().float().cpu().numpy().reshape(-1))
            mn = mn_np.ctypes.data_as(C.POINTER(C.c_float))
            mx = mx_np.ctypes.data_as(C.POINTER(C.c_float))
        with torch.cuda.device(self.device):
            stream = _lib.stream_ptr()  # the model's device's current stream, not the caller's current device's
            if self.handle is not None and structure == self._structure:
                _lib.check(self.lib.b200pets_model_refresh(self.handle, W, Bv, mem, nm, ns, mn, mx, stream), "model_refresh")
            else:
                self.close()
                nd = list(getattr(self.src, "no_delta_list", []) or [])
                nd_arr = (C.c_int32 * max(len(nd), 1))(*nd)
                h = C.c_void_p()
                _lib.check(self.lib.b200pets_model_create(C.byref(desc), W, Bv, mem, nm, ns, mn, mx, nd_arr, len(nd), stream,
                                                          C.byref(h)), "model_create")
                self.handle = h
                self._structure = structure
        self.desc = desc
        self._sig = sig

    def supports_tc(self) -> bool:
        return bool(self.lib.b200pets_model_supports_tc(self.handle))

    def close(self):
        if self.handle is not None:
            self.lib.b200pets_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
