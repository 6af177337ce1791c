"""TD3 behind the reference's ``oprl.algos.td3.TD3`` surface (td3.py:16-146)."""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass, field
from typing import Any

import torch as t
from torch import nn

from ..engine import EngineSpec
from .base_algorithm import EngineAdam, OffPolicyAlgorithm
from .nn_models import DeterministicPolicy, DoubleCritic


@dataclass
class TD3(OffPolicyAlgorithm):
    logger: Any
    state_dim: int
    action_dim: int
    batch_size: int = 256
    policy_noise: float = 0.2
    expl_noise: float = 0.1
    noise_clip: float = 0.5
    policy_freq: int = 2
    gamma: float = 0.99
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    max_action: float = 1.0
    tau: float = 5e-3
    log_every: int = 5000
    device: str = "cuda"

    actor: Any = field(init=False)
    actor_target: Any = field(init=False)
    optim_actor: Any = field(init=False)
    critic: nn.Module = field(init=False)
    critic_target: nn.Module = field(init=False)
    optim_critic: Any = field(init=False)
    update_step: int = 0
    _created: bool = False

    def create(self) -> "TD3":
        self.actor = DeterministicPolicy(
            state_dim=self.state_dim, action_dim=self.action_dim, hidden_units=(256, 256),
            hidden_activation=nn.ReLU(inplace=True), expl_noise=self.expl_noise, device=self.device)
        self.actor_target = deepcopy(self.actor).eval()
        self.critic = DoubleCritic(self.state_dim, self.action_dim, (256, 256), nn.ReLU(inplace=True))
        self.critic_target = deepcopy(self.critic).eval()
        self._start_engine(EngineSpec(
            algo="td3", state_dim=self.state_dim, action_dim=self.action_dim, n_critics=2,
            gamma=self.gamma, tau=self.tau, lr_actor=self.lr_actor, lr_critic=self.lr_critic,
            policy_noise=self.policy_noise, noise_clip=self.noise_clip, max_action=self.max_action))
        self.optim_actor = EngineAdam(self.engine, "actor", self.lr_actor, "step_actor")
        self.optim_critic = EngineAdam(self.engine, "critic", self.lr_critic, "step_critic")
        self._created = True
        return self

    def _wants_actor_step(self) -> bool:
        return self.update_step % self.policy_freq == 0  # td3.py:81

    def update(self, state: t.Tensor, action: t.Tensor, reward: t.Tensor, done: t.Tensor,
               next_state: t.Tensor) -> None:
        self._hand_batch(state, action, reward, done, next_state)
        actor_step = self.update_step % self.policy_freq == 0  # td3.py:81
        self._run_update(actor_step)
        if self.update_step % self.log_every == 0:  # td3.py:118-132,143-146
            sc = self.engine.scalars()
            self.logger.log_scalar("algo/q1", sc["q_mean"], self.update_step)
            self.logger.log_scalar("algo/q_target", sc["q_target_mean"], self.update_step)
            self.logger.log_scalar("algo/abs_q_err", sc["q_err_mean"], self.update_step)
            self.logger.log_scalar("algo/critic_loss", sc["critic_loss"], self.update_step)
            if actor_step:
                self.logger.log_scalar("algo/loss_actor", sc["actor_loss"], self.update_step)
        self.update_step += 1
