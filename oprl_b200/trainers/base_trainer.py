"""Single-process training loop around the GPU engine ("next" row N1 of SURVEY.md section 8f).

Same dataclass fields, call order and logging cadence as the reference's ``BaseTrainer``
(trainers/base_trainer.py:18-120): env step -> ``add_transition`` -> ``sample`` -> ``update`` ->
evaluate / save / log.  What changes is how the two hot calls meet the device:

* ``algo.attach_buffer(replay_buffer)`` fuses ``sample()`` and ``update()`` (the gather writes the
  GEMM operand layout; no copy-in),
* the sampled ``rewards`` are only pulled to the host on logging steps (the reference's
  ``rewards.mean().item()``, base_trainer.py:80, synchronises there as well),
* the policy acts on a CPU mirror of the actor refreshed by async D2H copies into pinned memory
  (``algo.enable_host_rollout``), and ``add_transition`` stages into a pinned ring flushed by one H2D +
  one scatter kernel: the environment loop never waits for the device,
* ``save_policy_every`` writes a CPU copy of the actor module that unpickles without the engine
  (scripts/visualize_policy_from_weights.py:66 only needs ``.exploit``).
"""
from __future__ import annotations

import copy
import logging
from dataclasses import dataclass
from typing import Any, Callable

import numpy as np
import torch as t

log = logging.getLogger(__name__)


@dataclass
class BaseTrainer:
    logger: Any
    env: Any
    make_env_test: Callable[[int], Any]
    replay_buffer: Any
    algo: Any
    gamma: float = 0.99
    num_steps: int = int(1e6)
    start_steps: int = int(10e3)
    batch_size: int = 128
    eval_interval: int = int(2e3)
    num_eval_episodes: int = 10
    save_buffer_every: int = 0
    save_policy_every: int = int(100_000)
    estimate_q_every: int = 0
    stdout_log_every: int = int(1e5)
    device: str = "cuda"
    seed: int = 0
    host_rollout_refresh_every: int = 1  # 0: act with the device network (H2D + sync per env step)

    def train(self) -> None:
        self.algo.check_created()
        self.replay_buffer.check_created()
        if hasattr(self.algo, "attach_buffer") and getattr(self.replay_buffer, "_engine", None) is None:
            self.algo.attach_buffer(self.replay_buffer)
        if self.host_rollout_refresh_every > 0 and hasattr(self.algo, "enable_host_rollout") \
                and getattr(self.algo, "_mirror", None) is None:
            self.algo.enable_host_rollout(self.host_rollout_refresh_every)

        state, _ = self.env.reset()
        for env_step in range(self.num_steps + 1):
            # rollout on the host (external simulator): uniform actions during warm-up
            if env_step <= self.start_steps:
                action = self.env.sample_action()
            else:
                action = self.algo.actor.explore(state)
            next_state, reward, terminated, truncated, _ = self.env.step(action)
            self.replay_buffer.add_transition(state, action, reward, terminated,
                                              episode_done=terminated or truncated)
            state = self.env.reset()[0] if (terminated or truncated) else next_state

            if len(self.replay_buffer) < self.batch_size:
                continue
            batch = self.replay_buffer.sample(self.batch_size)
            self.algo.update(*batch)

            if env_step % self.eval_interval == 0:
                self._log_evaluation(env_step, batch[2])
            if self.save_policy_every > 0 and env_step % self.save_policy_every == 0:
                self._save_policy(env_step)
            if self.stdout_log_every > 0 and env_step % self.stdout_log_every == 0:
                log.info("env step %d, sampled reward mean %.4f", env_step, float(batch[2].mean()))

    def _log_evaluation(self, env_step: int, rewards: t.Tensor) -> None:
        metrics = self.evaluate()
        rb = self.replay_buffer
        for tag, value in (("trainer/ep_reward", metrics["return"]),
                           ("trainer/avg_reward", float(rewards.mean())),
                           ("trainer/buffer_transitions", len(rb)),
                           ("trainer/buffer_episodes", rb.episodes_counter),
                           ("trainer/buffer_last_ep_len", rb.last_episode_length)):
            self.logger.log_scalar(tag, value, env_step)

    def evaluate(self) -> dict[str, float]:
        mirror = getattr(self.algo, "_mirror", None)
        if mirror is not None:
            mirror.refresh_now()  # evaluate the current weights, not a copy one update old
        returns = []
        for episode in range(self.num_eval_episodes):
            env = self.make_env_test(self.seed + episode)
            state, _ = env.reset()
            total, done = 0.0, False
            while not done:
                state, reward, terminated, truncated, _ = env.step(self.algo.actor.exploit(state))
                total += reward
                done = terminated or truncated
            returns.append(total)
        return {"return": float(np.mean(returns))}

    def _save_policy(self, env_step: int) -> None:
        path = self.logger.log_dir / "weights" / f"{env_step}.w"
        path.parent.mkdir(parents=True, exist_ok=True)
        t.save(export_policy(self.algo.actor), path)


def export_policy(actor: t.nn.Module) -> t.nn.Module:
    """Stand-alone CPU copy of the actor (fresh tensors, not views into the engine arena): what
    ``t.save(self.algo.actor, ...)`` produces in the reference (base_trainer.py:113-120)."""
    clone = copy.deepcopy(actor)
    with t.no_grad():
        for p in clone.parameters():
            p.data = p.data.detach().cpu().clone()
    for attr in ("_device", "device"):
        if hasattr(clone, attr):
            setattr(clone, attr, "cpu")
    clone._load_state_dict_post_hooks.clear()
    return clone
