"""The reference's own functional tests, run against the oprl_b200 classes:
tests/functional/test_rl_algos.py:17-57 (all four algorithms + SAC with and without temperature tuning:
exploit / explore return 1-d actions, then one update() on a batch of 8 with an int64 `done` and next_state
aliasing state) and tests/functional/test_replay_buffer.py:5-21.  The environment is stubbed by its
dimensions (walker-walk: 24 observations, 6 actions) -- dm_control is outside the hot path.

Plus the drop-in properties the reference's trainers rely on: `torch.save(algo.actor)`
(trainers/base_trainer.py `_save_policy`), and update() training on exactly the tensors it is passed."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OBS_DIM, ACT_DIM = 24, 6  # DMControlEnv("walker-walk") observation / action sizes


class NullLogger:
    log_dir = "/tmp"

    def log_scalar(self, *a, **k):
        pass

    def log_scalars(self, *a, **k):
        pass


def algo_classes():
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.sac import SAC
    from oprl_b200.algos.td3 import TD3
    from oprl_b200.algos.tqc import TQC

    return {"DDPG": DDPG, "SAC": SAC, "TD3": TD3, "TQC": TQC}


def _run_common_test(algo, obs):
    # body of the reference's _run_common_test (test_rl_algos.py:17-31), unchanged but for the env stub
    action = algo.actor.exploit(obs)
    assert action.ndim == 1

    action = algo.actor.explore(obs)
    assert action.ndim == 1

    _batch_size = 8
    batch_obs = torch.randn(_batch_size, OBS_DIM)
    batch_actions = torch.clamp(torch.randn(_batch_size, ACT_DIM), -1, 1)
    batch_rewards = torch.randn(_batch_size, 1)
    batch_dones = torch.randint(2, (_batch_size, 1))
    before = {k: v["theta"].clone() for k, v in algo.engine.arena.items()}
    algo.update(batch_obs, batch_actions, batch_rewards, batch_dones, batch_obs)
    torch.cuda.synchronize()
    for k, v in algo.engine.arena.items():
        assert torch.isfinite(v["theta"]).all(), k
        assert not torch.equal(v["theta"], before[k]), f"{k} parameters did not move"


@pytest.mark.parametrize("name", ["DDPG", "SAC", "TD3", "TQC"])
def test_ddpg_td3_tqc(name):
    torch.manual_seed(0)
    obs = np.random.default_rng(0).standard_normal(OBS_DIM).astype(np.float32)
    algo = algo_classes()[name](logger=NullLogger(), state_dim=OBS_DIM, action_dim=ACT_DIM, device="cuda").create()
    _run_common_test(algo, obs)


@pytest.mark.parametrize("tune_alpha", [True, False])
def test_sac(tune_alpha):
    torch.manual_seed(0)
    obs = np.random.default_rng(1).standard_normal(OBS_DIM).astype(np.float32)
    algo = algo_classes()["SAC"](logger=NullLogger(), tune_alpha=tune_alpha, state_dim=OBS_DIM, action_dim=ACT_DIM,
                                 device="cuda").create()
    _run_common_test(algo, obs)


def test_replay_buffer():
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer
    from oprl_b200.buffers.protocols import ReplayBufferProtocol

    state_dim = 7
    max_episode_length = 10
    num_transitions = 100
    buffer = EpisodicReplayBuffer(
        buffer_size_transitions=num_transitions,
        state_dim=state_dim,
        action_dim=3,
        max_episode_lenth=max_episode_length,
    ).create()
    assert isinstance(buffer, ReplayBufferProtocol)

    states = buffer.states
    assert len(states.shape) == 3
    assert states.shape[0] == num_transitions // max_episode_length
    assert states.shape[1] == max_episode_length + 1
    assert states.shape[2] == state_dim


@pytest.mark.parametrize("name", ["DDPG", "SAC"])
def test_policy_pickles_like_the_reference_trainer_saves_it(name):
    """`t.save(self.algo.actor, path)` of the reference's BaseTrainer._save_policy / distrib save_policy must work on
    an engine-backed actor (its load_state_dict hook is a picklable object, not a closure) and load without an engine."""
    algo = algo_classes()[name](logger=NullLogger(), state_dim=OBS_DIM, action_dim=ACT_DIM, device="cuda").create()
    buf = io.BytesIO()
    torch.save(algo.actor, buf)
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)  # (as scripts/visualize_policy_from_weights.py does: same device)
    for (k, a), (k2, b) in zip(algo.actor.state_dict().items(), loaded.state_dict().items()):
        assert k == k2 and torch.equal(a.cpu(), b.cpu()), k
    loaded.load_state_dict(algo.get_policy_state_dict())  # the detached hook is a no-op
    obs = np.zeros(OBS_DIM, np.float32)
    assert loaded.exploit(obs).shape == (ACT_DIM,)


def _fresh_ddpg_with_buffer(seed=0):
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    torch.manual_seed(seed)
    algo = DDPG(logger=NullLogger(), state_dim=OBS_DIM, action_dim=ACT_DIM, device="cuda").create()
    buf = EpisodicReplayBuffer(buffer_size_transitions=2000, state_dim=OBS_DIM, action_dim=ACT_DIM, max_episode_lenth=100,
                               device="cuda").create()
    g = torch.Generator(device="cuda").manual_seed(1)
    E = 10
    buf.states[:E, :100].normal_(generator=g)
    buf.actions[:E].uniform_(-1, 1, generator=g)
    buf.rewards[:E].uniform_(0, 1, generator=g)
    for e in range(E):
        buf.ep_lens[e] = 100
    buf._number_transitions = E * 100
    buf._ep_pointer = E
    buf.episodes_counter = E + 1
    algo.attach_buffer(buf)
    return algo, buf


def _theta(algo):
    torch.cuda.synchronize()
    return {k: v["theta"].clone() for k, v in algo.engine.arena.items()}


def test_update_trains_on_the_tensors_it_is_passed():
    """the in-place fast path must not silently substitute the engine's own gathered batch.
    (1) an older batch kept across a later sample(); (2) a replaced tensor (r * 2); (3) an in-place edit."""
    B = 32
    # reference behaviour for each case = a fresh learner (same init) updated with explicit copies
    def expect(batch):
        algo, _ = _fresh_ddpg_with_buffer()
        algo.update(*[x.clone() for x in batch])
        return _theta(algo)

    # (1) b1 survives a later sample() and update(*b1) uses b1
    algo, buf = _fresh_ddpg_with_buffer()
    np.random.seed(3)
    b1 = buf.sample(B)
    keep = [x.clone() for x in b1]
    b2 = buf.sample(B)
    assert not torch.equal(b1[0], b2[0])
    for x, k in zip(b1, keep):
        assert torch.equal(x, k), "a later sample() overwrote an earlier batch"
    algo.update(*b1)
    want = expect(keep)
    for k, v in _theta(algo).items():
        assert torch.equal(v, want[k]), k

    # (2) a replaced tensor
    algo, buf = _fresh_ddpg_with_buffer()
    np.random.seed(3)
    s, a, r, d, s2 = buf.sample(B)
    algo.update(s, a, r * 2.0, d, s2)
    want = expect([s, a, r * 2.0, d, s2])
    for k, v in _theta(algo).items():
        assert torch.equal(v, want[k]), k

    # (3) an in-place edit after sampling
    algo, buf = _fresh_ddpg_with_buffer()
    np.random.seed(3)
    batch = buf.sample(B)
    batch[2].mul_(0.5)
    algo.update(*batch)
    want = expect(batch)
    for k, v in _theta(algo).items():
        assert torch.equal(v, want[k]), k

    # and the untouched hand-off still matches the explicit path bit for bit
    algo, buf = _fresh_ddpg_with_buffer()
    np.random.seed(3)
    batch = buf.sample(B)
    algo.update(*batch)
    want = expect(batch)
    for k, v in _theta(algo).items():
        assert torch.equal(v, want[k]), k
