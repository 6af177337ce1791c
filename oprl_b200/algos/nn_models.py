"""Host-side network containers (rollout, initialisation, state_dict / pickle format).

Mirrors the public surface of the reference's ``oprl.algos.nn_models`` (nn_models.py:27-214):
same class names, constructor arguments, attribute names and ``state_dict`` keys
(``mlp.nn.{0,2,4}.*``, ``q1.nn.*``, ``net.nn.*``) so checkpoints and env workers interoperate.
The gradient update does NOT run through these modules: their parameters are views into the
engine's flat fp32 arena (see ``adopt_parameters``) and the CUDA engine updates them in place.
"""
from __future__ import annotations

import math
from typing import Final, Iterable

import numpy as np
import numpy.typing as npt
import torch as t
import torch.nn as nn
from torch.nn.functional import logsigmoid

LOG_STD_MIN_MAX: Final[tuple[float, float]] = (-20, 2)


def initialize_weight_orthogonal(m: nn.Module, gain: float = nn.init.calculate_gain("relu")) -> None:
    """Orthogonal weights (gain sqrt(2)), zero bias -- reference nn_models.py:14-17."""
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data, gain)
        m.bias.data.zero_()


class MLP(nn.Module):
    """Linear/activation stack; ``self.nn`` is the Sequential the reference exposes."""

    def __init__(self, input_dim: int, output_dim: int, hidden_units: tuple[int, ...] = (64, 64),
                 hidden_activation: nn.Module = nn.Tanh(),
                 output_activation: nn.Module = nn.Identity()) -> None:
        super().__init__()
        widths = [input_dim, *hidden_units]
        mods: list[nn.Module] = []
        for a, b in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(a, b), hidden_activation]
        mods += [nn.Linear(widths[-1], output_dim), output_activation]
        self.nn = nn.Sequential(*mods)

    def forward(self, x: t.Tensor) -> t.Tensor:
        return self.nn(x)


class Critic(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_units: tuple[int, ...] = (256, 256),
                 hidden_activation: nn.Module = nn.ReLU(inplace=True)) -> None:
        super().__init__()
        self.q1 = MLP(state_dim + action_dim, 1, hidden_units, hidden_activation)

    def forward(self, states: t.Tensor, actions: t.Tensor) -> t.Tensor:
        return self.q1(t.cat([states, actions], dim=-1))

    Q1 = forward


class DoubleCritic(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_units: tuple[int, ...] = (256, 256),
                 hidden_activation: nn.Module = nn.ReLU(inplace=True)) -> None:
        super().__init__()
        self.q1 = MLP(state_dim + action_dim, 1, hidden_units, hidden_activation)
        self.q2 = MLP(state_dim + action_dim, 1, hidden_units, hidden_activation)

    def forward(self, states: t.Tensor, actions: t.Tensor) -> tuple[t.Tensor, t.Tensor]:
        x = t.cat([states, actions], dim=-1)
        return self.q1(x), self.q2(x)

    def Q1(self, states: t.Tensor, actions: t.Tensor) -> t.Tensor:
        return self.q1(t.cat([states, actions], dim=-1))


class QuantileQritic(nn.Module):
    """n_nets quantile critics ``qf{i}`` of 3x512 hidden units (reference tqc.py:39-58)."""

    def __init__(self, state_dim: int, action_dim: int, n_quantiles: int, n_nets: int) -> None:
        super().__init__()
        self.n_quantiles = n_quantiles
        self.n_nets = n_nets
        self.nets = []
        for i in range(n_nets):
            net = MLP(state_dim + action_dim, n_quantiles, (512, 512, 512), hidden_activation=nn.ReLU())
            self.add_module(f"qf{i}", net)
            self.nets.append(net)

    def forward(self, state: t.Tensor, action: t.Tensor) -> t.Tensor:
        sa = t.cat((state, action), dim=1)
        return t.stack([net(sa) for net in self.nets], dim=1)


class DeterministicPolicy(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_units: tuple[int, ...] = (256, 256),
                 hidden_activation: nn.Module = nn.ReLU(inplace=True), max_action: float = 1.0,
                 expl_noise: float = 0.1, device: str = "cpu") -> None:
        super().__init__()
        self.mlp = MLP(state_dim, action_dim, hidden_units, hidden_activation).apply(
            initialize_weight_orthogonal)
        self._device = device
        self._action_shape = action_dim
        self._max_action = max_action
        self._expl_noise = expl_noise

    def forward(self, states: t.Tensor) -> t.Tensor:
        return t.tanh(self.mlp(states))

    def exploit(self, state: npt.NDArray) -> npt.NDArray:
        mirror = self.__dict__.get("_host_mirror")
        if mirror is not None:  # rollouts on the host copy (algos/host_mirror.py): no device round trip
            return mirror.exploit(state)
        with t.no_grad():
            x = t.as_tensor(state).unsqueeze(0).to(self._device)
            return self.forward(x).cpu().numpy().flatten()

    def explore(self, state: npt.NDArray) -> npt.NDArray:
        mirror = self.__dict__.get("_host_mirror")
        if mirror is not None:
            return mirror.explore(state)
        # reference quirk kept: exploration acts on the pre-tanh output (nn_models.py:144-150)
        with t.no_grad():
            x = t.as_tensor(state, device=self._device).unsqueeze(0)
            noise = (t.randn(self._action_shape) * self._expl_noise).to(self._device)
            action = (self.mlp(x) + noise).cpu().numpy()[0]
        return np.clip(action, -self._max_action, self._max_action)


class TanhNormal:
    """tanh-squashed diagonal Gaussian (reference nn_models.py:197-214)."""

    def __init__(self, normal_mean: t.Tensor, normal_std: t.Tensor, device: str) -> None:
        self.normal_mean = normal_mean
        self.normal_std = normal_std

    def log_prob(self, pre_tanh: t.Tensor) -> t.Tensor:
        z = (pre_tanh - self.normal_mean) / self.normal_std
        gauss = -0.5 * z * z - self.normal_std.log() - 0.5 * math.log(2 * math.pi)
        return gauss - (2 * math.log(2) + logsigmoid(2 * pre_tanh) + logsigmoid(-2 * pre_tanh))

    def rsample(self) -> tuple[t.Tensor, t.Tensor]:
        pre = self.normal_mean + self.normal_std * t.randn_like(self.normal_mean)
        return t.tanh(pre), pre


class GaussianActor(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_units: tuple[int, ...],
                 hidden_activation: nn.Module, device: str) -> None:
        super().__init__()
        self.action_dim = action_dim
        self.net = MLP(state_dim, 2 * action_dim, hidden_units, hidden_activation=hidden_activation)
        self.device = device

    def forward(self, obs: t.Tensor) -> tuple[t.Tensor, t.Tensor | None]:
        mean, log_std = self.net(obs).split([self.action_dim, self.action_dim], dim=1)
        if not self.training:
            return t.tanh(mean), None
        dist = TanhNormal(mean, log_std.clamp(*LOG_STD_MIN_MAX).exp(), self.device)
        action, pre = dist.rsample()
        return action, dist.log_prob(pre).sum(dim=1, keepdim=True)

    def explore(self, state: npt.NDArray) -> npt.NDArray:
        mirror = self.__dict__.get("_host_mirror")
        if mirror is not None:  # rollouts on the host copy (algos/host_mirror.py): no device round trip
            return mirror.explore(state)
        with t.no_grad():
            action, _ = self.forward(t.as_tensor(state, device=self.device).unsqueeze(0))
        return action.cpu().numpy()[0]

    def exploit(self, state: npt.NDArray) -> npt.NDArray:
        mirror = self.__dict__.get("_host_mirror")
        if mirror is not None:
            return mirror.exploit(state)
        was_training = self.training
        self.eval()
        try:
            return self.explore(state)
        finally:
            self.train(was_training)


def adopt_parameters(flat: t.Tensor, modules: Iterable[nn.Module]) -> None:
    """Move every parameter of ``modules`` (in ``parameters()`` order, the arena layout of
    oprl_engine_arena_floats) into ``flat`` and rebind ``param.data`` to the arena view, so the
    engine's in-place Adam / Polyak updates are what ``state_dict()`` and rollouts see."""
    off = 0
    with t.no_grad():
        for mod in modules:
            for p in mod.parameters():
                n = p.numel()
                view = flat[off:off + n].view(p.shape)
                view.copy_(p.data.to(flat.device))
                p.data = view
                off += n
    if off != flat.numel():
        raise ValueError(f"arena holds {flat.numel()} floats but the modules have {off}")
