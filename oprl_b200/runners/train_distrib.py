"""Spawn ``num_env_workers`` rollout processes and the learner process(es) (reference
runners/train_distrib.py:14-40) around shared-memory queues.  ``config.num_learners`` (default 1; an
oprl_b200 extension for the 8 actors -> 8 GPU learners configuration) spawns one data-parallel learner per GPU."""
from __future__ import annotations

import multiprocessing as mp
import socket
from typing import Any, Callable

from ..distrib.queue import QueueServer


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_distrib_training(run_env_worker: Callable, run_policy_update_worker: Callable,
                         make_env: Callable[[int], Any], make_algo: Callable[[Any], Any],
                         make_policy: Callable[[], Any], make_replay_buffer: Callable[[], Any],
                         make_logger: Callable[[], Any], config: Any) -> None:
    ctx = mp.get_context("spawn")  # the learner initialises CUDA: never fork
    n_learners = int(getattr(config, "num_learners", 1) or 1)
    names = [f"{kind}_{i}" for i in range(config.num_env_workers) for kind in ("env", "policy")]
    with QueueServer(names):
        procs = [ctx.Process(target=run_env_worker, args=(make_env, make_policy, config, i))
                 for i in range(config.num_env_workers)]
        if n_learners == 1:
            procs.append(ctx.Process(target=run_policy_update_worker,
                                     args=(make_algo, make_env, make_replay_buffer, make_logger, config)))
        else:
            port = _free_port()
            for r in range(n_learners):
                procs.append(ctx.Process(target=run_policy_update_worker,
                                         args=(make_algo, make_env, make_replay_buffer, make_logger, config,
                                               r, n_learners, port)))
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        bad = [p.exitcode for p in procs if p.exitcode]
        if bad:
            raise RuntimeError(f"distributed training: worker exit codes {bad}")
