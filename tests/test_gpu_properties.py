"""GPU: size-independent properties at BASELINE.json's full sizes (1e6-transition replay, batch 256):
bit-exact gather, run-to-run determinism, and the Adam / Polyak update rules checked in place."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class NullLogger:
    log_dir = "/tmp"

    def log_scalar(self, *a, **k):
        pass

    def log_scalars(self, *a, **k):
        pass


def full_buffer(S, A, episodes=1000, seed=0):
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    buf = EpisodicReplayBuffer(buffer_size_transitions=1_000_000, state_dim=S, action_dim=A).create()
    g = torch.Generator(device="cuda").manual_seed(seed)
    buf.states[:episodes, :1000].normal_(generator=g)
    buf.actions[:episodes].uniform_(-1, 1, generator=g)
    buf.rewards[:episodes].uniform_(0, 1, generator=g)
    for e in range(episodes):
        buf.ep_lens[e] = 1000
    buf._number_transitions = episodes * 1000
    buf._ep_pointer = 0
    buf.episodes_counter = episodes
    return buf


@pytest.mark.parametrize("attached", [False, True])
def test_gather_is_bit_exact_at_full_buffer_size(attached):
    from oprl_b200.algos.ddpg import DDPG

    S, A, B = 24, 6, 256
    buf = full_buffer(S, A)
    if attached:
        algo = DDPG(logger=NullLogger(), state_dim=S, action_dim=A).create()
        algo.attach_buffer(buf)
    np.random.seed(3)
    ep_step = buf.draw_indices(B)
    # include the edges: first / last transition of the first / last episode
    ep_step[:4] = [[0, 0], [0, 999], [999, 0], [999, 999]]
    np.random.seed(3)
    _ = np.random.randint(0, 1, 1)
    out = buf._engine.sample(B, ep_step) if attached else buf.gather(ep_step)
    ep = torch.from_numpy(ep_step[:, 0].astype(np.int64)).cuda()
    st = torch.from_numpy(ep_step[:, 1].astype(np.int64)).cuda()
    ref = (buf.states[ep, st], buf.actions[ep, st], buf.rewards[ep, st], buf.dones[ep, st], buf.states[ep, st + 1])
    for got, want in zip(out, ref):
        assert torch.equal(got, want)


def make_pair(cls, **kw):
    torch.manual_seed(0)
    a = cls(logger=NullLogger(), state_dim=24, action_dim=6, **kw).create()
    b = cls(logger=NullLogger(), state_dim=24, action_dim=6, **kw).create()
    for grp in ("actor", "critic"):
        for key in ("theta", "target"):
            if a.engine.arena[grp][key] is not None:
                b.engine.arena[grp][key].copy_(a.engine.arena[grp][key])
    b.engine.mark_params_dirty()
    return a, b


@pytest.mark.parametrize("name", ["ddpg", "td3", "sac"])
def test_device_resident_learner_is_deterministic(name):
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.sac import SAC
    from oprl_b200.algos.td3 import TD3

    cls, kw = dict(ddpg=(DDPG, {}), td3=(TD3, {}), sac=(SAC, dict(tune_alpha=True)))[name]
    a, b = make_pair(cls, **kw)
    buf = full_buffer(24, 6, episodes=100)
    for algo in (a, b):
        algo.attach_buffer(buf)
        algo.engine.set_prefix(buf.ep_lens[:buf.episodes_counter])
    for _ in range(40):
        a.learner_step(256)
    for _ in range(40):
        b.learner_step(256)
    torch.cuda.synchronize()
    for grp in ("actor", "critic"):
        for key in ("theta", "m", "v", "target"):
            x, y = a.engine.arena[grp][key], b.engine.arena[grp][key]
            if x is not None:
                assert torch.equal(x, y), (grp, key)
                assert torch.isfinite(x).all()
    assert a.engine.state().log_alpha == b.engine.state().log_alpha
    assert a.engine.state().tick == 40 and a.engine.state().step_critic == 40
    assert a.engine.state().step_actor == (20 if name == "td3" else 40)


def test_first_update_obeys_adam_and_polyak_rules_exactly():
    """After one DDPG update at full size: m = lerp(0, g, 1-b1), v = (1-b2) g g, target = tau*theta +
    (1-tau)*target_old, theta = theta_old - lr/(1-b1) * m / (sqrt(v)/sqrt(1-b2) + eps) -- evaluated with
    torch's own fp32 operators on the engine's gradient arena (bit-exact for m, v, target)."""
    from oprl_b200.algos.ddpg import DDPG

    torch.manual_seed(1)
    algo = DDPG(logger=NullLogger(), state_dim=24, action_dim=6).create()
    buf = full_buffer(24, 6, episodes=100)
    algo.attach_buffer(buf)
    algo.engine.set_prefix(buf.ep_lens[:buf.episodes_counter])
    ar = algo.engine.arena
    before = {g: {k: v.clone() for k, v in a.items() if v is not None} for g, a in ar.items()}
    algo.learner_step(256)
    torch.cuda.synchronize()
    for grp, lr in (("critic", 3e-4), ("actor", 3e-4)):
        n = before[grp]["theta"].numel()
        # the reference runs torch's CPU operators: evaluate the rule there (1 ulp slack for fma choices)
        g = ar[grp]["grad"][:n].cpu()
        m = torch.zeros_like(g).lerp_(g, 1 - 0.9)
        v = torch.zeros_like(g).mul_(0.999).addcmul_(g, g, value=1 - 0.999)
        torch.testing.assert_close(ar[grp]["m"].cpu(), m, rtol=1.3e-7, atol=0)
        torch.testing.assert_close(ar[grp]["v"].cpu(), v, rtol=1.3e-7, atol=0)
        denom = (v.sqrt() / (1 - 0.999) ** 0.5).add_(1e-8)
        theta = before[grp]["theta"].cpu().addcdiv_(m, denom, value=-lr / (1 - 0.9))
        assert (ar[grp]["theta"].cpu() - theta).abs().max().item() <= 1.6e-8  # 1 ulp at |theta| < 0.25
        target = 5e-3 * ar[grp]["theta"].cpu() + (1 - 5e-3) * before[grp]["target"].cpu()
        assert torch.equal(ar[grp]["target"].cpu(), target), grp
        assert g.abs().max().item() > 0


def test_pipelined_scalar_readback_matches_the_blocking_one():
    """scalars_async() of update t, consumed after updates t+1.. were launched, returns exactly what a
    blocking scalars() read right after update t returns (host-batch path, as bench.py's e2e loop)."""
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.engine import EngineSpec  # noqa: F401

    a, b = make_pair(DDPG)
    g = torch.Generator().manual_seed(7)
    batches = []
    for _ in range(6):
        batches.append([torch.randn(256, 24, generator=g).pin_memory(), (torch.rand(256, 6, generator=g) * 2 - 1).pin_memory(),
                        torch.rand(256, 1, generator=g).pin_memory(), torch.zeros(256, 1).pin_memory(),
                        torch.randn(256, 24, generator=g).pin_memory()])
    blocking = []
    for bt in batches:
        a.update(*bt)
        blocking.append(a.engine.scalars())
    pend = []
    for bt in batches:
        b.update(*bt)
        pend.append(b.engine.scalars_async())
    for want, p in zip(blocking, pend):
        got = p.result()
        assert got == want
    for grp in ("actor", "critic"):
        assert torch.equal(a.engine.arena[grp]["theta"], b.engine.arena[grp]["theta"])
    # tickets expire once the ring has wrapped (8 further updates, each with its read-back)
    old = b.engine.scalars_async()
    for _ in range(8):
        b.update(*batches[0])
        b.engine.scalars_async()
    with pytest.raises(Exception):
        old.result()


@pytest.mark.parametrize("name", ["ddpg", "td3", "sac"])
def test_prefetching_learner_step_equals_the_in_order_sequence(name):
    """learner_step() (oprl_step: after four steady steps the NEXT batch is gathered as a parallel branch
    of the running update graph, into the other working set) must produce bit-for-bit what the in-order
    API sequence sample(device draw) ; update() produces: same draws, same batches, same parameters --
    including across a replay change announced by set_prefix in the middle."""
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.sac import SAC
    from oprl_b200.algos.td3 import TD3

    cls, kw = dict(ddpg=(DDPG, {}), td3=(TD3, {}), sac=(SAC, dict(tune_alpha=True)))[name]
    a, b = make_pair(cls, **kw)
    buf = full_buffer(24, 6, episodes=100)
    for algo in (a, b):
        algo.attach_buffer(buf)
        algo.engine.set_prefix(buf.ep_lens[:buf.episodes_counter])
    B = 256
    for k in range(14):
        if k == 9:  # the replay binding changes: both must re-order behind it and sample from 60 episodes
            for algo in (a, b):
                algo.engine.set_prefix(buf.ep_lens[:60])
        a.learner_step(B)
        b.engine.sample(B, None)
        b._run_update(b._wants_actor_step())
        b._after_update()
    torch.cuda.synchronize()
    assert a.engine.state().tick == b.engine.state().tick == 14
    for grp in ("actor", "critic"):
        for key in ("theta", "m", "v", "target"):
            x, y = a.engine.arena[grp][key], b.engine.arena[grp][key]
            if x is not None:
                assert torch.equal(x, y), (grp, key)
    assert a.engine.state().log_alpha == b.engine.state().log_alpha


def test_rebuilt_programs_release_their_workspaces():
    """every program rebuild (new batch arena, world-size change) used to leak the previous program's
    activation workspaces until engine destroy.  Programs now own and free them."""
    from tests.test_gpu_parity import make_algo
    from tests.util import load_case

    fx = load_case("ddpg_b8")
    algo = make_algo(fx)
    eng = algo.engine
    g = torch.Generator().manual_seed(0)

    def one_update(B):
        batch = [torch.randn(B, 24, generator=g), torch.rand(B, 6, generator=g) * 2 - 1, torch.rand(B, 1, generator=g),
                 torch.zeros(B, 1), torch.randn(B, 24, generator=g)]
        algo.update(*[x.cuda() for x in batch])

    for B in (64, 64):
        one_update(B)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for k in range(12):
        eng.set_world_size(2 if k % 2 == 0 else 1)  # clears the programs: the next update rebuilds them
        one_update(64)
    eng.set_world_size(1)
    one_update(64)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    # one program's workspaces are ~2 MB; twelve leaked rebuilds would be ~25 MB
    assert free0 - free1 < 8 << 20, f"device memory shrank by {(free0 - free1) / 2**20:.1f} MiB over 12 program rebuilds"


def test_nstep_gather_matches_the_oracle_bit_for_bit():
    """n-step return assembly in the gather kernel (extension, oprl_buffer_set_nstep) against oracle.nstep_batch:
    windows cut by a done flag and by the episode end, n = 1 identical to the plain gather, and the update consumes
    the assembled batch (the effective done makes the 1-step target the n-step target)."""
    from oracle import oprl_oracle as O
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    class NullLogger:
        log_dir = "/tmp"

        def log_scalar(self, *a, **k):
            pass

        def log_scalars(self, *a, **k):
            pass

    S, A, Lmax, E = 24, 6, 40, 25
    rng = np.random.default_rng(5)
    algo = DDPG(logger=NullLogger(), state_dim=S, action_dim=A, device="cuda").create()
    buf = EpisodicReplayBuffer(buffer_size_transitions=E * Lmax, state_dim=S, action_dim=A, max_episode_lenth=Lmax, gamma=0.97).create()
    st = rng.standard_normal((E, Lmax + 1, S)).astype(np.float32)
    ac = rng.uniform(-1, 1, (E, Lmax, A)).astype(np.float32)
    rw = rng.uniform(0, 1, (E, Lmax, 1)).astype(np.float32)
    dn = (rng.uniform(0, 1, (E, Lmax, 1)) < 0.1).astype(np.float32)
    buf.states.copy_(torch.from_numpy(st))
    buf.actions.copy_(torch.from_numpy(ac))
    buf.rewards.copy_(torch.from_numpy(rw))
    buf.dones.copy_(torch.from_numpy(dn))
    lens = [int(x) for x in rng.integers(3, Lmax + 1, size=E - 1)]
    for e, n in enumerate(lens):
        buf.ep_lens[e] = n
    buf._number_transitions = sum(lens)
    buf._ep_pointer = E - 1
    buf.episodes_counter = E - 1
    algo.attach_buffer(buf)
    for n_step in (1, 3, 5):
        buf.n_step = n_step
        np.random.seed(11)
        ep_step = buf.draw_indices(200)
        np.random.seed(11)
        got = [x.cpu().numpy() for x in buf.sample(200)]
        want = O.nstep_batch(st, ac, rw, dn, buf.ep_lens, ep_step[:, 0], ep_step[:, 1], n_step, buf.gamma)
        for nm, g_, w_ in zip(("s", "a", "r", "d", "s2"), got, want):
            assert np.array_equal(g_, w_), (n_step, nm, np.abs(g_ - w_).max())
        if n_step > 1:
            assert (got[3] != dn[ep_step[:, 0], ep_step[:, 1]]).any()  # windows longer than one step were assembled
    algo.update(*buf.sample(64))
    torch.cuda.synchronize()
    assert np.isfinite(algo.engine.scalars()["critic_loss"])
    # the device-resident learner loop draws its indices on the GPU and assembles the same windows
    algo.engine.set_prefix(buf.ep_lens[:buf.episodes_counter])
    for _ in range(8):
        algo.learner_step(64)
    torch.cuda.synchronize()
    assert np.isfinite(algo.engine.scalars()["critic_loss"])
