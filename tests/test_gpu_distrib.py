"""GPU: actor/learner split over the shared-memory queue transport (2 CPU rollout workers -> 1 GPU learner;
with >= 2 GPUs also 2 workers -> 2 data-parallel learners, the scaled-down shape of BASELINE config 5)."""
import pathlib
import pickle
import tempfile

import numpy as np
import pytest
import torch

from tests.test_gpu_trainer import ToyEnv

pytestmark = pytest.mark.gpu
OUT = pathlib.Path(tempfile.gettempdir()) / "oprl_b200_distrib_test"


def make_env(seed):
    return ToyEnv(seed, horizon=25)


def make_policy():
    from oprl_b200.algos.nn_models import DeterministicPolicy

    return DeterministicPolicy(24, 6, device="cpu")


class FileLogger:
    log_dir = OUT

    def log_scalar(self, tag, value, step):
        pass

    def log_scalars(self, values, step):
        pass


def make_logger():
    return FileLogger()


def make_algo(logger):
    import os

    from oprl_b200.algos.ddpg import DDPG

    rank = int(os.environ.get("RANK", "0"))
    algo = DDPG(logger=logger, state_dim=24, action_dim=6, device=f"cuda:{rank}").create()
    real_update = algo.update
    counter = {"n": 0}

    def counted(*batch):
        counter["n"] += 1
        real_update(*batch)
        (OUT / f"updates{rank if rank else ''}.txt").write_text(str(counter["n"]))
        if counter["n"] % 50 == 0:  # replicas must stay bit-identical: keep a parameter checksum per rank
            import torch

            torch.cuda.synchronize()
            th = algo.engine.arena["actor"]["theta"]
            (OUT / f"theta_sum{rank}.txt").write_text(repr(float(th.double().sum())))

    algo.update = counted
    return algo


def make_buffer():
    import os

    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    return EpisodicReplayBuffer(buffer_size_transitions=5000, state_dim=24, action_dim=6,
                                max_episode_lenth=25, device=f"cuda:{os.environ.get('RANK', '0')}").create()


@pytest.mark.timeout(150)
def test_two_workers_feed_one_learner():
    from oprl_b200.distrib.env_worker import run_env_worker
    from oprl_b200.distrib.policy_update_worker import run_policy_update_worker
    from oprl_b200.runners.config import DistribConfig
    from oprl_b200.runners.train_distrib import run_distrib_training

    OUT.mkdir(exist_ok=True)
    (OUT / "updates.txt").write_text("0")
    cfg = DistribConfig(batch_size=32, num_env_workers=2, episodes_per_worker=4, warmup_epochs=0,
                        episode_length=25, learner_num_waits=4, warmup_env_steps=30)
    run_distrib_training(run_env_worker, run_policy_update_worker, make_env, make_algo, make_policy,
                         make_buffer, make_logger, cfg)
    # (the learner only counts a wait when a 1 s poll comes back empty AFTER the workers are up:
    # give the three child interpreters time to import torch on a cold box)
    # epochs 1..3 run episode_length * num_env_workers = 50 updates each
    assert int((OUT / "updates.txt").read_text()) == 150


@pytest.mark.timeout(240)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_workers_feed_two_data_parallel_learners():
    from oprl_b200.distrib.env_worker import run_env_worker
    from oprl_b200.distrib.policy_update_worker import run_policy_update_worker
    from oprl_b200.runners.config import DistribConfig
    from oprl_b200.runners.train_distrib import run_distrib_training

    OUT.mkdir(exist_ok=True)
    for name in ("updates.txt", "updates1.txt"):
        (OUT / name).write_text("0")
    cfg = DistribConfig(batch_size=32, num_env_workers=2, episodes_per_worker=4, warmup_epochs=0,
                        episode_length=25, learner_num_waits=4, warmup_env_steps=30, num_learners=2)
    run_distrib_training(run_env_worker, run_policy_update_worker, make_env, make_algo, make_policy,
                         make_buffer, make_logger, cfg)
    assert int((OUT / "updates.txt").read_text()) == 150
    assert int((OUT / "updates1.txt").read_text()) == 150
    # both replicas applied the same all-reduced gradients: identical parameters
    assert (OUT / "theta_sum0.txt").read_text() == (OUT / "theta_sum1.txt").read_text()
