"""GPU: the trainer loop, buffer ingest, the sample()/update() fusion and checkpoint/resume."""
import pathlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class ToyEnv:
    """Deterministic stand-in for a dm_control task (never terminates, truncates at `horizon`)."""

    def __init__(self, seed, S=24, A=6, horizon=25):
        self.rng = np.random.default_rng(seed)
        self.S, self.A, self.horizon, self.t = S, A, horizon, 0

    def reset(self):
        self.t = 0
        self.state = self.rng.standard_normal(self.S).astype(np.float32)
        return self.state, {}

    def sample_action(self):
        return self.rng.uniform(-1, 1, self.A)  # float64, like dm_control (environment/dm_control.py:38)

    def step(self, action):
        self.t += 1
        self.state = (0.9 * self.state + 0.1 * self.rng.standard_normal(self.S)).astype(np.float32)
        reward = float(-np.square(action).sum() * 0.01 + self.state[0])
        return self.state, reward, False, self.t >= self.horizon, {}


class MemLogger:
    def __init__(self, tmp):
        self.log_dir = pathlib.Path(tmp)
        self.rows = []

    def log_scalar(self, tag, value, step):
        self.rows.append((tag, float(value), step))

    def log_scalars(self, values, step):
        for k, v in values.items():
            self.log_scalar(k, v, step)


def test_buffer_ingest_and_sample_match_the_reference_fixture():
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer
    from tests.util import load_case

    fx = load_case("buffer")
    buf = EpisodicReplayBuffer(buffer_size_transitions=int(fx["cap"]), state_dim=int(fx["S"]),
                               action_dim=int(fx["A"]), max_episode_lenth=int(fx["L"])).create()
    for s, a, r, d, e in zip(fx["log_s"], fx["log_a"], fx["log_r"], fx["log_d"], fx["log_ep_done"]):
        buf.add_transition(s, a, float(r), bool(d), episode_done=bool(e))
    assert list(buf.ep_lens) == list(fx["ep_lens"])
    assert buf._ep_pointer == int(fx["ep_pointer"]) and buf.episodes_counter == int(fx["episodes_counter"])
    assert len(buf) == int(fx["n_transitions"])
    # every row the reference wrote is bit-identical (rows it never wrote are unspecified there)
    for name in ("states", "actions", "rewards", "dones"):
        assert np.array_equal(getattr(buf, name).cpu().numpy(), fx[name]), name
    np.random.seed(0)
    out = buf.sample(64)
    for nm, x in zip(("s", "a", "r", "d", "s2"), out):
        assert np.array_equal(x.cpu().numpy(), fx["batch_" + nm]), nm


@pytest.mark.parametrize("algo_name", ["ddpg", "td3", "sac", "tqc"])
def test_trainer_runs_and_resumes(tmp_path, algo_name):
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.sac import SAC
    from oprl_b200.algos.td3 import TD3
    from oprl_b200.algos.tqc import TQC
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer
    from oprl_b200.checkpoint import load_checkpoint, save_checkpoint
    from oprl_b200.trainers.base_trainer import BaseTrainer, export_policy

    cls = dict(ddpg=DDPG, td3=TD3, sac=SAC, tqc=TQC)[algo_name]
    kw = dict(tune_alpha=True) if algo_name == "sac" else {}
    logger = MemLogger(tmp_path)
    torch.manual_seed(0)
    np.random.seed(0)
    algo = cls(logger=logger, state_dim=24, action_dim=6, **kw).create()
    buf = EpisodicReplayBuffer(buffer_size_transitions=2000, state_dim=24, action_dim=6,
                               max_episode_lenth=25).create()
    trainer = BaseTrainer(logger=logger, env=ToyEnv(0), make_env_test=lambda s: ToyEnv(100 + s),
                          replay_buffer=buf, algo=algo, num_steps=150, start_steps=60, batch_size=32,
                          eval_interval=50, num_eval_episodes=2, save_policy_every=100, stdout_log_every=0)
    before = algo.engine.arena["actor"]["theta"].clone()
    trainer.train()
    assert len(buf) == 151 and buf.episodes_counter == 7
    assert not torch.equal(before, algo.engine.arena["actor"]["theta"])
    assert torch.isfinite(algo.engine.arena["critic"]["theta"]).all()
    tags = {r[0] for r in logger.rows}
    assert {"trainer/ep_reward", "trainer/avg_reward", "trainer/buffer_transitions"} <= tags
    # saved policy unpickles to a CPU module with .exploit (scripts/visualize_policy_from_weights.py)
    pol = torch.load(tmp_path / "weights" / "100.w", weights_only=False)
    assert pol.exploit(np.zeros(24, np.float32)).shape == (6,)
    assert all(p.device.type == "cpu" for p in pol.parameters())
    live = export_policy(algo.actor)
    obs = np.ones(24, np.float32)
    if getattr(algo, "_mirror", None) is not None:
        algo._mirror.refresh_now()  # the host mirror may still be one asynchronous copy behind the last update
    np.testing.assert_allclose(live.exploit(obs), algo.actor.exploit(obs), atol=1e-5)

    # checkpoint -> two more updates -> restore -> the same two updates reproduce bit-exactly
    ckpt = tmp_path / "full.ckpt"
    save_checkpoint(ckpt, algo, buf)
    np.random.seed(123)
    idx = [buf.draw_indices(32) for _ in range(2)]

    def two_updates():
        for k, ep_step in enumerate(idx):
            batch = buf.gather(ep_step)
            if algo_name != "ddpg":
                g = torch.Generator(device="cuda").manual_seed(7 + k)
                for which in range(1 if algo_name == "td3" else 2):
                    algo.engine.set_noise(which, torch.randn(32, 6, device="cuda", generator=g))
            algo.update(*batch)
        torch.cuda.synchronize()
        return {g: {k: v.clone() for k, v in a.items() if v is not None and k != "grad"}
                for g, a in algo.engine.arena.items()}, algo.engine.state().log_alpha

    ref, ref_alpha = two_updates()
    load_checkpoint(ckpt, algo, buf)
    got, got_alpha = two_updates()
    for g in ref:
        for k in ref[g]:
            assert torch.equal(ref[g][k], got[g][k]), (g, k)
    assert ref_alpha == got_alpha


def test_staged_ingest_wraps_the_pinned_ring_and_stays_bit_exact():
    """add_transition stages into a pinned ring (4096 rows) flushed by one H2D + one scatter launch: more
    transitions than the ring holds, flushes forced at odd points, episode ring wrap -- storage must equal a
    plain numpy replay of the same bookkeeping."""
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    S, A, Lmax, E = 5, 3, 50, 120
    buf = EpisodicReplayBuffer(buffer_size_transitions=E * Lmax, state_dim=S, action_dim=A, max_episode_lenth=Lmax).create()
    rng = np.random.default_rng(0)
    ref = {"states": np.zeros((E, Lmax + 1, S), np.float32), "actions": np.zeros((E, Lmax, A), np.float32),
           "rewards": np.zeros((E, Lmax, 1), np.float32), "dones": np.zeros((E, Lmax, 1), np.float32)}
    ep, lens = 0, [0] * E
    n = 9500  # > 2 ring capacities, > E * Lmax: the episode ring wraps as well
    for i in range(n):
        s, a = rng.standard_normal(S).astype(np.float32), rng.uniform(-1, 1, A).astype(np.float32)
        r, d = float(rng.uniform()), bool(rng.uniform() < 0.01)
        done_ep = d or lens[ep] == Lmax - 1
        buf.add_transition(s, a, r, d, episode_done=done_ep)
        ref["states"][ep, lens[ep]] = s
        ref["actions"][ep, lens[ep]] = a
        ref["rewards"][ep, lens[ep], 0] = np.float32(r)
        ref["dones"][ep, lens[ep], 0] = float(d)
        lens[ep] += 1
        if done_ep:
            ep = (ep + 1) % E
            lens[ep] = 0
        if i % 1777 == 5:
            buf.flush()
    assert list(buf.ep_lens) == lens and buf._ep_pointer == ep
    for name in ("states", "actions", "rewards", "dones"):
        got = getattr(buf, name).cpu().numpy()
        # rows of episodes that were reset keep their old content in both (only ep_lens shrinks)
        assert np.array_equal(got, ref[name]), name


def test_host_rollout_mirror_follows_the_device_weights_without_syncing():
    """explore / exploit on a CPU mirror refreshed by async D2H copies into pinned memory."""
    import time

    from oprl_b200.algos.ddpg import DDPG

    class NullLogger:
        log_dir = "/tmp"

        def log_scalar(self, *a, **k):
            pass

        def log_scalars(self, *a, **k):
            pass

    torch.manual_seed(0)
    algo = DDPG(logger=NullLogger(), state_dim=24, action_dim=6).create()
    obs = np.random.default_rng(0).standard_normal(24).astype(np.float32)
    want = algo.actor.exploit(obs)  # device path
    # env-steps/s of the acting call alone (no-op environment), device path vs host mirror
    def rate(fn, n=300):
        fn(obs)
        t0 = time.perf_counter()
        for _ in range(n):
            fn(obs)
        return n / (time.perf_counter() - t0)

    dev_rate = rate(algo.actor.explore)
    mirror = algo.enable_host_rollout(refresh_every=1)
    assert mirror._np is not None, "the numpy fast path must cover the reference's policy classes"
    np.testing.assert_allclose(algo.actor.exploit(obs), want, atol=1e-6)
    host_rate = rate(algo.actor.explore)
    print(f"explore(): {dev_rate:.0f} calls/s on the device network, {host_rate:.0f} calls/s on the host mirror ({host_rate / dev_rate:.1f}x)")
    assert host_rate > 2.0 * dev_rate
    # the mirror follows the updates
    g = torch.Generator().manual_seed(1)
    batch = [torch.randn(64, 24, generator=g), torch.rand(64, 6, generator=g) * 2 - 1, torch.rand(64, 1, generator=g),
             torch.zeros(64, 1), torch.randn(64, 24, generator=g)]
    for _ in range(5):
        algo.update(*batch)
    torch.cuda.synchronize()
    algo.actor.exploit(obs)  # picks up a completed copy
    assert mirror.swaps >= 1
    mirror.refresh_now()
    algo.disable_host_rollout()
    dev = algo.actor.exploit(obs)
    algo.actor.__dict__["_host_mirror"] = mirror
    np.testing.assert_allclose(algo.actor.exploit(obs), dev, atol=1e-5)
    assert not np.allclose(dev, want)  # the weights did move
    # and the policy still pickles with the mirror attached (torch.save(algo.actor) of the reference trainer)
    import io

    bio = io.BytesIO()
    torch.save(algo.actor, bio)
