"""TQC behind the reference's ``oprl.algos.tqc.TQC`` surface (tqc.py:60-189): n_nets quantile
critics, drop of the top atoms, quantile-Huber loss, always-on temperature tuning."""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch as t
import torch.nn as nn

from ..engine import EngineSpec
from .base_algorithm import EngineAdam, OffPolicyAlgorithm
from .nn_models import GaussianActor, QuantileQritic  # noqa: F401  (QuantileQritic re-exported)


def quantile_huber_loss_f(quantiles: t.Tensor, samples: t.Tensor, device: str) -> t.Tensor:
    """Host version of the loss the engine's tqc_loss_kernel computes (tqc.py:14-36);
    quantiles [B, nets, nq], samples [B, kept]."""
    delta = samples[:, None, None, :] - quantiles[:, :, :, None]
    mag = delta.abs()
    huber = t.where(mag > 1, mag - 0.5, 0.5 * delta * delta)
    nq = quantiles.shape[2]
    tau = (t.arange(nq, device=device, dtype=t.float32) + 0.5) / nq
    return ((tau[None, None, :, None] - (delta < 0).float()).abs() * huber).mean()


@dataclass
class TQC(OffPolicyAlgorithm):
    logger: Any
    state_dim: int
    action_dim: int
    gamma: float = 0.99
    lr_actor = 3e-4   # un-annotated on purpose: not constructor kwargs in the reference either
    lr_critic = 3e-4  # (tqc.py:67-69)
    lr_alpha = 3e-4
    tau: float = 0.005
    top_quantiles_to_drop: int = 2
    n_quantiles: int = 25
    n_nets: int = 5
    log_every: int = 5000
    device: str = "cuda"

    actor: Any = field(init=False)
    actor_target: Any = field(init=False, default=None)
    actor_optimizer: Any = field(init=False)
    critic: QuantileQritic = field(init=False)
    critic_target: QuantileQritic = field(init=False)
    critic_optimizer: Any = field(init=False)
    target_entropy: float = field(init=False)
    alpha_optimizer: Any = field(init=False, default=None)
    quantiles_total: int = field(init=False)
    update_step: int = 0
    _created: bool = False

    def create(self) -> "TQC":
        self.target_entropy = -np.prod(self.action_dim).item()
        self.actor = GaussianActor(self.state_dim, self.action_dim, hidden_units=(256, 256),
                                   hidden_activation=nn.ReLU(), device=self.device)
        self.critic = QuantileQritic(self.state_dim, self.action_dim, self.n_quantiles, self.n_nets)
        self.critic_target = copy.deepcopy(self.critic)
        self.quantiles_total = self.n_quantiles * self.n_nets
        self._start_engine(EngineSpec(
            algo="tqc", state_dim=self.state_dim, action_dim=self.action_dim,
            critic_hidden=512, critic_layers=3, n_critics=self.n_nets, n_quantiles=self.n_quantiles,
            top_quantiles_to_drop=self.top_quantiles_to_drop, tune_alpha=True, gamma=self.gamma,
            tau=self.tau, lr_actor=self.lr_actor, lr_critic=self.lr_critic, lr_alpha=self.lr_alpha,
            alpha_init=0.2, target_entropy=float(self.target_entropy)))
        self.actor_optimizer = EngineAdam(self.engine, "actor", self.lr_actor, "step_actor")
        self.critic_optimizer = EngineAdam(self.engine, "critic", self.lr_critic, "step_critic")
        self._created = True
        return self

    @property
    def log_alpha(self) -> t.Tensor:
        return t.tensor(self.engine.state().log_alpha, dtype=t.float64)

    def update(self, state: t.Tensor, action: t.Tensor, reward: t.Tensor, done: t.Tensor,
               next_state: t.Tensor) -> None:
        self._hand_batch(state, action, reward, done, next_state)
        self._run_update(True)
        if self.update_step % self.log_every == 0:  # tqc.py:179-187
            sc = self.engine.scalars()
            self.logger.log_scalars({"algo/critic_loss": sc["critic_loss"],
                                     "algo/actor_loss": sc["actor_loss"],
                                     "algo/alpha_loss": sc["alpha_loss"]}, self.update_step)
        self.update_step += 1
