"""SAC behind the reference's ``oprl.algos.sac.SAC`` surface (sac.py:16-155)."""
from __future__ import annotations

import math
from copy import deepcopy
from dataclasses import dataclass, field
from typing import Any

import torch as t
from torch import nn

from ..engine import EngineSpec
from .base_algorithm import EngineAdam, OffPolicyAlgorithm
from .nn_models import DoubleCritic, GaussianActor


@dataclass
class SAC(OffPolicyAlgorithm):
    logger: Any
    state_dim: int
    action_dim: int
    batch_size: int = 256
    tune_alpha: bool = False
    gamma: float = 0.99
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    lr_alpha: float = 1e-3
    alpha_init: float = 0.2
    target_update_coef: float = 5e-3
    device: str = "cuda"
    log_every: int = 5000

    actor: Any = field(init=False)
    actor_target: Any = field(init=False, default=None)
    optim_actor: Any = field(init=False)
    critic: nn.Module = field(init=False)
    critic_target: nn.Module = field(init=False)
    optim_critic: Any = field(init=False)
    update_step: int = 0
    _created: bool = False

    def create(self) -> "SAC":
        self.actor = GaussianActor(self.state_dim, self.action_dim, (256, 256), nn.ReLU(inplace=True),
                                   device=self.device)
        self.critic = DoubleCritic(self.state_dim, self.action_dim, (256, 256), nn.ReLU(inplace=True))
        self.critic_target = deepcopy(self.critic).eval()
        self.target_entropy = -float(self.action_dim)
        self._start_engine(EngineSpec(
            algo="sac", state_dim=self.state_dim, action_dim=self.action_dim, n_critics=2,
            tune_alpha=self.tune_alpha, gamma=self.gamma, tau=self.target_update_coef,
            lr_actor=self.lr_actor, lr_critic=self.lr_critic, lr_alpha=self.lr_alpha,
            alpha_init=self.alpha_init, target_entropy=self.target_entropy))
        self.optim_actor = EngineAdam(self.engine, "actor", self.lr_actor, "step_actor")
        self.optim_critic = EngineAdam(self.engine, "critic", self.lr_critic, "step_critic")
        self._created = True
        return self

    # the temperature lives on the device (float64 log_alpha + its Adam state); reading it
    # synchronises, so only logging / checkpoint code should touch these.
    @property
    def log_alpha(self) -> t.Tensor:
        return t.tensor(self.engine.state().log_alpha, dtype=t.float64)

    @property
    def alpha(self) -> float:
        return math.exp(self.engine.state().log_alpha) if self.tune_alpha else self.alpha_init

    def update(self, state: t.Tensor, action: t.Tensor, reward: t.Tensor, done: t.Tensor,
               next_state: t.Tensor) -> None:
        self._hand_batch(state, action, reward, done, next_state)
        self._run_update(True)
        if self.update_step % self.log_every == 0:  # sac.py:112-121,143-155
            sc = self.engine.scalars()
            self.logger.log_scalars({"algo/q1": sc["q_mean"], "algo/q_target": sc["q_target_mean"],
                                     "algo/abs_q_err": sc["q_err_mean"],
                                     "algo/critic_loss": sc["critic_loss"]}, self.update_step)
            if self.tune_alpha:
                self.logger.log_scalar("algo/loss_alpha", sc["alpha_loss"], self.update_step)
            self.logger.log_scalars({"algo/loss_actor": sc["actor_loss"], "algo/alpha": sc["alpha"],
                                     "algo/log_pi": sc["logpi_mean"]}, self.update_step)
        self.update_step += 1
