"""Spawn ``num_env_workers`` rollout processes and one learner process (reference
runners/train_distrib.py:14-40) around the in-node queue server."""
from __future__ import annotations

import multiprocessing as mp
from typing import Any, Callable

from ..distrib.queue import QueueServer


def run_distrib_training(run_env_worker: Callable, run_policy_update_worker: Callable,
                         make_env: Callable[[int], Any], make_algo: Callable[[Any], Any],
                         make_policy: Callable[[], Any], make_replay_buffer: Callable[[], Any],
                         make_logger: Callable[[], Any], config: Any) -> None:
    ctx = mp.get_context("spawn")  # the learner initialises CUDA: never fork
    with QueueServer():
        procs = [ctx.Process(target=run_env_worker, args=(make_env, make_policy, config, i))
                 for i in range(config.num_env_workers)]
        procs.append(ctx.Process(target=run_policy_update_worker,
                                 args=(make_algo, make_env, make_replay_buffer, make_logger, config)))
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        bad = [p.exitcode for p in procs if p.exitcode]
        if bad:
            raise RuntimeError(f"distributed training: worker exit codes {bad}")
