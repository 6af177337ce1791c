"""DDPG behind the reference's ``oprl.algos.ddpg.DDPG`` surface (ddpg.py:17-107): same
dataclass fields, ``create()``, ``update(state, action, reward, done, next_state)``; the
update itself is one CUDA-graph launch of the sm_100a engine."""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass, field
from typing import Any

import torch as t
from torch import nn

from ..engine import EngineSpec
from .base_algorithm import EngineAdam, OffPolicyAlgorithm
from .nn_models import Critic, DeterministicPolicy


@dataclass
class DDPG(OffPolicyAlgorithm):
    logger: Any
    state_dim: int
    action_dim: int
    expl_noise: float = 0.1
    gamma: float = 0.99
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    tau: float = 5e-3
    batch_size: int = 256
    max_action: float = 1.0
    device: str = "cuda"

    actor: Any = field(init=False)
    actor_target: Any = field(init=False)
    optim_actor: Any = field(init=False)
    critic: nn.Module = field(init=False)
    critic_target: nn.Module = field(init=False)
    optim_critic: Any = field(init=False)
    update_step: int = 0
    _created: bool = False

    def create(self) -> "DDPG":
        self.actor = DeterministicPolicy(
            state_dim=self.state_dim, action_dim=self.action_dim, hidden_units=(256, 256),
            hidden_activation=nn.ReLU(inplace=True), expl_noise=self.expl_noise,
            max_action=self.max_action, device=self.device)
        self.actor_target = deepcopy(self.actor)
        self.critic = Critic(self.state_dim, self.action_dim)
        self.critic_target = deepcopy(self.critic)
        self._start_engine(EngineSpec(
            algo="ddpg", state_dim=self.state_dim, action_dim=self.action_dim, n_critics=1,
            gamma=self.gamma, tau=self.tau, lr_actor=self.lr_actor, lr_critic=self.lr_critic,
            max_action=self.max_action))
        self.optim_actor = EngineAdam(self.engine, "actor", self.lr_actor, "step_actor")
        self.optim_critic = EngineAdam(self.engine, "critic", self.lr_critic, "step_critic")
        self._created = True
        return self

    def _after_update(self) -> None:
        pass  # reference quirk kept: DDPG never advances update_step

    def update(self, state: t.Tensor, action: t.Tensor, reward: t.Tensor, done: t.Tensor,
               next_state: t.Tensor) -> None:
        self._hand_batch(state, action, reward, done, next_state)
        self._run_update(True)  # reference quirk kept: update_step never advances
