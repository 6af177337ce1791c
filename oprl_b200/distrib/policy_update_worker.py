"""Learner(s): one episode per worker per epoch into the GPU replay buffer, then
``episode_length * num_env_workers`` gradient updates on the engine, then the new actor weights
to every worker (reference distrib/policy_update_worker.py:22-119).

``num_learners > 1`` (configs/distrib_ddpg.py:24-31 scaled out: N CPU actors -> N GPU learners): one learner
process per GPU, replicated replay buffer and parameters.  Rank 0 receives the episodes from the workers'
shared-memory rings and broadcasts each one ONCE over NCCL / NVLink; every rank ingests it with one H2D-free
scatter kernel, draws its own minibatch rows and runs the update with the gradient all-reduce fused into the
Adam kernels (``enable_data_parallel``).  Rank 0 alone talks to the workers, evaluates and saves.
"""
from __future__ import annotations

import logging
import os
from itertools import count
from typing import Any, Callable

import numpy as np
import torch as t

from ..trainers.base_trainer import export_policy
from .queue import Queue, episode_rows_to_list, pack_weights, unpack_episode

log = logging.getLogger(__name__)


def _ingest_rows(buffer, rows: np.ndarray, S: int, A: int) -> None:
    buffer.add_episode(episode_rows_to_list(rows, S, A))


def run_policy_update_worker(make_algo: Callable[[Any], Any], make_env_test: Callable[[int], Any],
                             make_buffer: Callable[[], Any], make_logger: Callable[[], Any],
                             config: Any, rank: int = 0, world: int = 1, master_port: int = 0) -> None:
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(master_port), RANK=str(rank),
                          WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        t.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=t.device(f"cuda:{rank}"))
    algo = make_algo(make_logger())
    buffer = make_buffer()
    if hasattr(algo, "attach_buffer"):
        algo.attach_buffer(buffer)  # sample() gathers straight into the GEMM operand layout
    if world > 1:
        algo.enable_data_parallel()
        np.random.seed(1234 + rank)  # every replica draws its own rows of the global minibatch
    dev = algo.engine.device
    lead = rank == 0
    q_envs = [Queue(f"env_{i}") for i in range(config.num_env_workers)] if lead else []
    q_policies = [Queue(f"policy_{i}") for i in range(config.num_env_workers)] if lead else []
    S, A = buffer.state_dim, buffer.action_dim
    W = 2 * S + A + 2

    for i_epoch in count(0):
        n_waits = 0
        for i_w in range(config.num_env_workers):
            rows = None
            if lead:
                data = None
                while data is None:
                    # the very first episodes include the workers' start-up (interpreter + simulator)
                    data = q_envs[i_w].pop_wait(1.0 if i_epoch else 30.0)
                    if data is None:
                        n_waits += 1
                        if n_waits == config.learner_num_waits:
                            log.info("learner is not receiving data, exiting")
                            break
                if data is not None:
                    rows, _, _ = unpack_episode(data)
            if world > 1:
                # one broadcast per episode: length first (0 = rank 0 gave up), then the rows
                n = t.tensor([0 if rows is None else rows.shape[0]], device=dev, dtype=t.int64)
                dist.broadcast(n, src=0)
                if int(n) == 0:
                    dist.destroy_process_group()
                    return
                payload = t.empty(int(n), W, device=dev, dtype=t.float32)
                if lead:
                    payload.copy_(t.from_numpy(np.ascontiguousarray(rows)))
                dist.broadcast(payload, src=0)
                rows = rows if lead else payload.cpu().numpy()
            elif rows is None:
                return
            _ingest_rows(buffer, rows, S, A)

        if i_epoch > config.warmup_epochs:
            for _ in range(config.episode_length * config.num_env_workers):
                algo.update(*buffer.sample(config.batch_size))

        if lead:
            weights = pack_weights(algo.get_policy_state_dict())
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
