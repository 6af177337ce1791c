"""Rollout worker: runs episodes on host CPU cores with a local copy of the policy, ships each
episode to the learner and picks up fresh weights (reference distrib/env_worker.py:15-65)."""
from __future__ import annotations

import logging
from typing import Any, Callable

import numpy as np

from .queue import Queue, pack_episode, unpack_weights

log = logging.getLogger(__name__)


def run_env_worker(make_env: Callable[[int], Any], make_policy: Callable[[], Any], config: Any,
                   id_worker: int) -> None:
    env = make_env(seed=0)
    policy = make_policy()
    q_env, q_policy = Queue(f"env_{id_worker}"), Queue(f"policy_{id_worker}")
    total_env_step = 0
    for i_ep in range(config.episodes_per_worker):
        episode = []
        state, _ = env.reset()
        for _ in range(config.episode_length):
            if total_env_step <= config.warmup_env_steps:
                action = env.sample_action()
            else:
                action = policy.explore(state)
            next_state, reward, terminated, truncated, _ = env.step(action)
            episode.append([state, action, reward, terminated, next_state])
            if terminated or truncated:
                break
            state = next_state
            total_env_step += 1
        s_dim, a_dim = int(np.asarray(episode[0][0]).size), int(np.asarray(episode[0][1]).size)
        q_env.push(pack_episode(episode, s_dim, a_dim))  # raw float32 rows through shared memory, no pickle
        # lock-step with the learner: wait for the weights it publishes after consuming the episode
        data, waited = None, 0.0
        while data is None:
            data = q_policy.pop_wait(2.0)
            if data is None:
                waited += 2.0
                log.info("worker %d waiting for the policy", id_worker)
                if waited >= 2.0 * 10 * config.learner_num_waits:
                    log.warning("worker %d: the learner is gone, exiting", id_worker)
                    return
        policy.load_state_dict(unpack_weights(data))
    log.info("env worker %d done", id_worker)
