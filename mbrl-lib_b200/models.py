"""Parameter containers with the attribute layout the staging code reads.

The drop-in use case hands *mbrl-lib's own* ``OneDTransitionRewardModel(GaussianMLP)`` objects to
:class:`mbrl_lib_b200.ModelEnv` (duck typed, SURVEY.md section 8b "Model (read-only by the fast path)").
These classes exist so that tests, ``bench.py`` and users without mbrl-lib installed can build the same
structure: ``.model.hidden_layers[i][0].{weight[E,K,N], bias[E,1,N]}``, ``.model.mean_and_logvar``,
``.model.{min,max}_logvar``, ``.model.elite_models``, ``.input_normalizer.{mean,std}`` ...
(mbrl/models/gaussian_mlp.py:86-127, mbrl/models/one_dim_tr_model.py:84-101).  They hold parameters only;
training stays in mbrl-lib / PyTorch (out of scope).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

_ACT = {"relu": nn.ReLU, "silu": nn.SiLU, "leaky_relu": lambda: nn.LeakyReLU(0.01)}


class EnsembleLinearLayer(nn.Module):
    """weight [E, in, out], bias [E, 1, out]   (layout of mbrl/models/util.py:41-45)."""

    def __init__(self, num_members: int, in_size: int, out_size: int):
        super().__init__()
        self.num_members, self.in_size, self.out_size = num_members, in_size, out_size
        self.weight = nn.Parameter(torch.zeros(num_members, in_size, out_size))
        self.bias = nn.Parameter(torch.zeros(num_members, 1, out_size))


class GaussianMLP(nn.Module):
    def __init__(self, in_size: int, out_size: int, device, num_layers: int = 4, ensemble_size: int = 1,
                 hid_size: int = 200, deterministic: bool = False, propagation_method: Optional[str] = None,
                 activation: str = "relu"):
        super().__init__()
        self.in_size, self.out_size, self.num_members = in_size, out_size, ensemble_size
        self.deterministic = deterministic
        self.propagation_method = propagation_method
        self.device = torch.device(device)
        layers = [nn.Sequential(EnsembleLinearLayer(ensemble_size, in_size, hid_size), _ACT[activation]())]
        for _ in range(num_layers - 1):
            layers.append(nn.Sequential(EnsembleLinearLayer(ensemble_size, hid_size, hid_size), _ACT[activation]()))
        self.hidden_layers = nn.Sequential(*layers)
        self.mean_and_logvar = EnsembleLinearLayer(ensemble_size, hid_size, out_size * (1 if deterministic else 2))
        if not deterministic:
            self.min_logvar = nn.Parameter(-10 * torch.ones(1, out_size), requires_grad=False)
            self.max_logvar = nn.Parameter(0.5 * torch.ones(1, out_size), requires_grad=False)
        self.elite_models: Optional[List[int]] = None
        self.to(self.device)

    def __len__(self):
        return self.num_members

    def set_elite(self, elite_indices: Sequence[int]):  # gaussian_mlp.py:377-379
        if len(elite_indices) != self.num_members:
            self.elite_models = list(elite_indices)

    def set_propagation_method(self, propagation_method: Optional[str] = None):
        self.propagation_method = propagation_method


class Normalizer:
    """mean / std of the model input, [1, in]   (mbrl/util/math.py:95-143)."""

    def __init__(self, in_size: int, device, dtype=torch.float32):
        self.mean = torch.zeros((1, in_size), device=device, dtype=dtype)
        self.std = torch.ones((1, in_size), device=device, dtype=dtype)
        self.eps = 1e-12 if dtype == torch.double else 1e-5
        self.device = device

    def update_stats(self, data):
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(data).to(self.device)
        self.mean = data.mean(0, keepdim=True)
        self.std = data.std(0, keepdim=True)
        self.std[self.std < self.eps] = 1.0


class OneDTransitionRewardModel:
    def __init__(self, model: GaussianMLP, target_is_delta: bool = True, normalize: bool = False,
                 normalize_double_precision: bool = False, learned_rewards: bool = True,
                 obs_process_fn: Optional[Callable] = None, no_delta_list: Optional[List[int]] = None,
                 num_elites: Optional[int] = None):
        self.model = model
        self.device = model.device
        self.input_normalizer: Optional[Normalizer] = None
        if normalize:
            self.input_normalizer = Normalizer(model.in_size, self.device,
                                               dtype=torch.double if normalize_double_precision else torch.float)
        self.learned_rewards = learned_rewards
        self.target_is_delta = target_is_delta
        self.no_delta_list = no_delta_list if no_delta_list else []
        self.obs_process_fn = obs_process_fn
        self.num_elites = num_elites or model.num_members

    def set_elite(self, elite_indices: Sequence[int]):
        self.model.set_elite(elite_indices)

    def set_propagation_method(self, propagation_method: Optional[str] = None):
        self.model.set_propagation_method(propagation_method)

    def __len__(self):
        return len(self.model)


def model_from_arrays(spec, arrays, device) -> OneDTransitionRewardModel:
    """Build the container for a ``synthetic.CaseSpec`` and its seeded arrays on ``device``."""
    from . import functions

    mlp = GaussianMLP(spec.in_size, spec.out_size, device, num_layers=spec.num_layers, ensemble_size=spec.ensemble_size,
                      hid_size=spec.hid_size, deterministic=spec.deterministic, propagation_method=spec.propagation,
                      activation=spec.activation)
    with torch.no_grad():
        for li, layer in enumerate(mlp.hidden_layers):
            layer[0].weight.copy_(torch.from_numpy(arrays["weights"][li]))
            layer[0].bias.copy_(torch.from_numpy(arrays["biases"][li]))
        mlp.mean_and_logvar.weight.copy_(torch.from_numpy(arrays["weights"][-1]))
        mlp.mean_and_logvar.bias.copy_(torch.from_numpy(arrays["biases"][-1]))
        if not spec.deterministic:
            mlp.min_logvar.copy_(torch.from_numpy(arrays["min_logvar"]))
            mlp.max_logvar.copy_(torch.from_numpy(arrays["max_logvar"]))
    wrapper = OneDTransitionRewardModel(
        mlp, target_is_delta=spec.target_is_delta, normalize=spec.normalize is not None,
        normalize_double_precision=spec.normalize == "float64", learned_rewards=spec.learned_rewards,
        obs_process_fn=functions.OBS_PROCESS_FNS.get(spec.obs_process), no_delta_list=list(spec.no_delta_list),
        num_elites=spec.num_models)
    if spec.normalize is not None:
        wrapper.input_normalizer.mean = torch.from_numpy(arrays["norm_mean"]).to(device)
        wrapper.input_normalizer.std = torch.from_numpy(arrays["norm_std"]).to(device)
    if spec.elites is not None:
        wrapper.set_elite(list(spec.elites))
    return wrapper
