"""Learner: one episode per worker per epoch into the GPU replay buffer, then
``episode_length * num_env_workers`` gradient updates on the engine, then the new actor weights
to every worker (reference distrib/policy_update_worker.py:22-119)."""
from __future__ import annotations

import logging
import pickle
from itertools import count
from typing import Any, Callable

import numpy as np
import torch as t

from ..trainers.base_trainer import export_policy
from .queue import Queue

log = logging.getLogger(__name__)


def run_policy_update_worker(make_algo: Callable[[Any], Any], make_env_test: Callable[[int], Any],
                             make_buffer: Callable[[], Any], make_logger: Callable[[], Any],
                             config: Any) -> None:
    algo = make_algo(make_logger())
    buffer = make_buffer()
    if hasattr(algo, "attach_buffer"):
        algo.attach_buffer(buffer)  # sample() gathers straight into the GEMM operand layout
    q_envs = [Queue(f"env_{i}") for i in range(config.num_env_workers)]
    q_policies = [Queue(f"policy_{i}") for i in range(config.num_env_workers)]

    for i_epoch in count(0):
        n_waits = 0
        for q in q_envs:
            data = None
            while data is None:
                # the very first episodes include the workers' start-up (interpreter + simulator)
                data = q.pop_wait(1.0 if i_epoch else 30.0)
                if data is None:
                    n_waits += 1
                    if n_waits == config.learner_num_waits:
                        log.info("learner is not receiving data, exiting")
                        return
            buffer.add_episode(pickle.loads(data))

        if i_epoch > config.warmup_epochs:
            for _ in range(config.episode_length * config.num_env_workers):
                algo.update(*buffer.sample(config.batch_size))

        weights = pickle.dumps({k: v.detach().cpu() for k, v in algo.get_policy_state_dict().items()})
        for q in q_policies:
            q.push(weights)

        if i_epoch > 0 and i_epoch % 10 == 0:
            mean_reward = evaluate(algo, make_env_test)
            algo.logger.log_scalar("trainer/ep_reward", mean_reward, i_epoch)
            save_policy(algo.actor, algo.logger.log_dir / "weights" / f"epoch_{i_epoch}.w")


def save_policy(policy, save_path) -> None:
    save_path.parent.mkdir(parents=True, exist_ok=True)
    t.save(export_policy(policy), save_path)


def evaluate(algo, make_env_test, num_eval_episodes: int = 5, seed: int = 0) -> float:
    returns = []
    for i_ep in range(num_eval_episodes):
        env = make_env_test(seed * 100 + i_ep)
        state, _ = env.reset()
        total, done = 0.0, False
        while not done:
            state, reward, terminated, truncated, _ = env.step(algo.actor.exploit(state))
            total += reward
            done = terminated or truncated
        returns.append(total)
    return float(np.mean(returns))
