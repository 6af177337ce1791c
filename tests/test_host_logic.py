"""CPU: host-side logic of the python shim -- index mapping, arena adoption, API surface -- and
the data-parallel scheme (world_size 2, gloo) checked with the CPU oracle as the compute."""
import dataclasses
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oprl_oracle as O
from tests.util import load_case


def test_inds_to_episodic_matches_reference_fixture_and_oracle():
    from oprl_b200.buffers.episodic_buffer import inds_to_episodic

    fx = load_case("buffer")
    ep, step = inds_to_episodic(fx["inds"], list(fx["ep_lens"]), int(fx["episodes_counter"]))
    assert (ep == fx["ep"]).all() and (step == fx["step"]).all()
    rng = np.random.default_rng(0)
    for _ in range(50):
        E = int(rng.integers(1, 12))
        lens = rng.integers(0, 9, size=E).tolist()
        if sum(lens) == 0:
            continue
        counter = int(rng.integers(1, E + 1))
        n = max(1, sum(lens[:counter]))
        inds = rng.integers(0, n, size=33)
        a = inds_to_episodic(inds.copy(), lens, counter)
        b = O.inds_to_episodic(inds.copy(), lens, counter)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_buffer_surface_matches_reference_protocol():
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer
    from oprl_b200.buffers.protocols import ReplayBufferProtocol

    buf = EpisodicReplayBuffer(buffer_size_transitions=100, state_dim=7, action_dim=3, max_episode_lenth=10)
    assert isinstance(buf, ReplayBufferProtocol)  # tests/functional/test_replay_buffer.py:17
    with pytest.raises(RuntimeError):
        buf.check_created()
    names = [f.name for f in dataclasses.fields(buf) if f.init]
    assert names == ["buffer_size_transitions", "state_dim", "action_dim", "gamma", "max_episode_lenth",
                     "episodes_counter", "device", "_ep_pointer", "_created"]  # episodic_buffer.py:14-27


def test_algorithm_constructor_fields_match_reference():
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.sac import SAC
    from oprl_b200.algos.td3 import TD3
    from oprl_b200.algos.tqc import TQC

    init = lambda cls: [f.name for f in dataclasses.fields(cls) if f.init]
    assert init(DDPG) == ["logger", "state_dim", "action_dim", "expl_noise", "gamma", "lr_actor", "lr_critic",
                          "tau", "batch_size", "max_action", "device", "update_step", "_created"]
    assert init(TD3)[:15] == ["logger", "state_dim", "action_dim", "batch_size", "policy_noise", "expl_noise",
                              "noise_clip", "policy_freq", "gamma", "lr_actor", "lr_critic", "max_action", "tau",
                              "log_every", "device"]
    assert init(SAC)[:13] == ["logger", "state_dim", "action_dim", "batch_size", "tune_alpha", "gamma", "lr_actor",
                              "lr_critic", "lr_alpha", "alpha_init", "target_update_coef", "device", "log_every"]
    assert init(TQC)[:10] == ["logger", "state_dim", "action_dim", "gamma", "tau", "top_quantiles_to_drop",
                              "n_quantiles", "n_nets", "log_every", "device"]
    algo = DDPG(logger=None, state_dim=3, action_dim=2)
    with pytest.raises(RuntimeError):
        algo.check_created()


def test_state_dict_keys_and_arena_adoption():
    from oprl_b200.algos.nn_models import (Critic, DeterministicPolicy, DoubleCritic, GaussianActor,
                                           QuantileQritic, adopt_parameters)

    pol = DeterministicPolicy(24, 6)
    assert list(pol.state_dict()) == [f"mlp.nn.{i}.{k}" for i in (0, 2, 4) for k in ("weight", "bias")]
    assert list(Critic(24, 6).state_dict())[0] == "q1.nn.0.weight"
    assert "q2.nn.4.bias" in DoubleCritic(24, 6).state_dict()
    assert "net.nn.4.weight" in GaussianActor(24, 6, (256, 256), torch.nn.ReLU(), "cpu").state_dict()
    assert "qf4.nn.6.bias" in QuantileQritic(24, 6, 25, 5).state_dict()
    n = sum(p.numel() for p in pol.parameters())
    assert n == 73_734
    flat = torch.zeros(n)
    before = torch.cat([p.detach().reshape(-1) for p in pol.parameters()])
    adopt_parameters(flat, [pol])
    assert torch.equal(flat, before)
    flat.mul_(2.0)  # the engine updates the arena in place: the module must see it
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in pol.parameters()]), before * 2)
    with pytest.raises(ValueError):
        adopt_parameters(torch.zeros(n + 1), [DeterministicPolicy(24, 6)])
    # orthogonal init with zero bias on the policy only (nn_models.py:14-17,128)
    w = DeterministicPolicy(24, 6).mlp.nn[2].weight
    assert torch.allclose(w @ w.t(), 2 * torch.eye(256), atol=1e-4)


def test_explore_exploit_shapes_on_cpu_modules():
    from oprl_b200.algos.nn_models import DeterministicPolicy, GaussianActor

    obs = np.random.default_rng(0).standard_normal(24).astype(np.float32)
    pol = DeterministicPolicy(24, 6)
    assert pol.exploit(obs).shape == (6,) and pol.explore(obs).shape == (6,)
    assert np.all(np.abs(pol.explore(obs)) <= 1.0)
    ga = GaussianActor(24, 6, (256, 256), torch.nn.ReLU(), "cpu")
    assert ga.exploit(obs).shape == (6,) and ga.explore(obs).shape == (6,)
    assert ga.training


# ------------------------------------------------------------------ data-parallel scheme
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    fx = load_case("ddpg")
    from tests.util import fixture_batch, oracle_from_fixture

    orc = oracle_from_fixture(fx)
    batch = fixture_batch(fx, 0)
    B = batch[0].shape[0]
    lo, hi = rank * B // world, (rank + 1) * B // world
    s, a, r, d, s2 = [x[lo:hi] for x in batch]
    sp = orc.spec
    # critic segment on this rank's rows, loss scaled by 1 / (world * B_local) as the engine does
    with torch.no_grad():
        y = r + (1.0 - d) * sp.gamma * O.critic_forward(orc.critics_target, s2, O.deterministic_policy(orc.actor_target, s2))[0]
    q = O.critic_forward(orc.critics, s, a)[0]
    loss = (q - y).pow(2).sum() / (world * (hi - lo))
    grads = torch.autograd.grad(loss, orc._critic_flat())
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)  # SUM, like the engine's NCCL all-reduce of the gradient arena
    if rank == 0:
        out.put(flat.numpy())
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_global_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fx = load_case("ddpg")
    from tests.util import oracle_from_fixture, run_fixture_updates

    orc = oracle_from_fixture(fx)
    run_fixture_updates(orc, fx, 0)
    ref = np.concatenate([g.reshape(-1).numpy() for g in orc.last_critic_grads])
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-7)
    np.testing.assert_allclose(ref, fx["first_critic_grad"], rtol=0, atol=1e-8)


def test_compat_overlay_redirects_reference_imports():
    import importlib
    import sys

    from oprl_b200 import compat

    saved = {k: sys.modules.get(k) for k in compat.OVERLAY}
    try:
        names = compat.install()
        assert set(names) == set(compat.OVERLAY)
        from oprl.algos.ddpg import DDPG  # the reference's import line (configs/ddpg.py:4)
        from oprl.buffers.episodic_buffer import EpisodicReplayBuffer

        assert DDPG.__module__ == "oprl_b200.algos.ddpg"
        assert EpisodicReplayBuffer.__module__ == "oprl_b200.buffers.episodic_buffer"
        assert importlib.import_module("oprl.algos.tqc").TQC.__module__ == "oprl_b200.algos.tqc"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_in_node_queue_surface():
    """distrib/queue.py:4-19 surface (push / pop -> bytes | None) on the broker-less shared-memory transport
    (more in tests/test_queue_transport.py)."""
    from oprl_b200.distrib.queue import Queue, QueueServer

    with QueueServer(["env_0", "policy_0"]):
        a, b = Queue("env_0"), Queue("env_0")
        assert b.pop() is None
        a.push(b"\x01\x02payload")
        assert b.pop_wait(1.0) == b"\x01\x02payload"
        assert b.pop_wait(0.05) is None
        assert Queue("policy_0").pop() is None
        a.close()
        b.close()


def test_config_records_match_reference_defaults():
    from oprl_b200.runners.config import CommonParameters, DistribConfig

    d = DistribConfig()
    assert (d.batch_size, d.num_env_workers, d.episodes_per_worker, d.warmup_epochs, d.episode_length,
            d.learner_num_waits, d.warmup_env_steps) == (128, 4, 100, 16, 1000, 10, 1000)
    c = CommonParameters(state_dim=24, action_dim=6, num_steps=100)
    assert (c.eval_every, c.estimate_q_every, c.log_every) == (2500, 5000, 2500)


class _FakeEngine:
    """Records the call order of the python data-parallel wrapper (no GPU, no CUDA library)."""

    def __init__(self):
        self.calls = []
        self.fused_comm = False
        self.arena = {"actor": {"grad": torch.ones(4), "theta": torch.zeros(4), "target": None},
                      "critic": {"grad": torch.ones(4) * 2, "theta": torch.zeros(4), "target": torch.zeros(4)}}

    def update(self, actor_step=True, segment=-1):
        self.calls.append(("update", actor_step, segment))

    def set_world_size(self, n):
        self.calls.append(("world", n))

    def mark_params_dirty(self):
        pass


def _segment_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oprl_b200.algos.base_algorithm import OffPolicyAlgorithm

    algo = OffPolicyAlgorithm()
    algo.engine = _FakeEngine()
    algo.engine.arena["critic"]["theta"] += rank  # replicas start different: rank 0's values must win
    algo.enable_data_parallel(fused=False)
    algo._run_update(True)
    algo._run_update(False)
    if rank == 0:
        out.put((algo.engine.calls, algo.engine.arena["critic"]["grad"].tolist(),
                 algo.engine.arena["actor"]["grad"].tolist(), algo.engine.arena["critic"]["theta"].tolist()))
    else:
        out.put(("theta", algo.engine.arena["critic"]["theta"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_style_wrapper_orders_segments_and_allreduces_over_gloo():
    from oprl_b200._lib import SEG_ACTOR_STEP, SEG_CRITIC_GRAD, SEG_CRITIC_STEP_ACTOR_GRAD

    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_segment_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    main = next(g for g in got if g[0] != "theta")
    other = next(g for g in got if g[0] == "theta")
    calls, cgrad, agrad, theta0 = main
    assert calls == [("world", 2),
                     ("update", True, SEG_CRITIC_GRAD), ("update", True, SEG_CRITIC_STEP_ACTOR_GRAD),
                     ("update", True, SEG_ACTOR_STEP),
                     ("update", False, SEG_CRITIC_GRAD), ("update", False, SEG_CRITIC_STEP_ACTOR_GRAD)]
    assert cgrad == [8.0] * 4  # 2 -> all-reduce(sum) over 2 ranks -> 4 -> again in the second update -> 8
    assert agrad == [2.0] * 4  # only the update with an actor step reduces the actor arena
    assert other[1] == theta0 == [0.0] * 4  # parameters broadcast from rank 0


def test_gemm_planning_code_on_the_host():
    """tools/host_logic_test.cu: CT32 index map, TMEM accumulator plan, split-K choice and the
    shared-memory budget of the GEMM kernel, compiled with nvcc and run on the CPU."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as g

    g.build()
    out = subprocess.run([os.path.join(root, "build", "host_logic_test")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout
