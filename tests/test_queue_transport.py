"""CPU: the shared-memory queue transport (oprl_b200/distrib/queue.py) -- the reference's Queue surface
(distrib/queue.py:4-19: push / pop) without a broker, plus the raw-float episode / weights formats."""
import multiprocessing as mp

import numpy as np
import torch

from oprl_b200.distrib.queue import (Queue, QueueServer, episode_rows_to_list, pack_episode, pack_weights,
                                     unpack_episode, unpack_weights)


def _producer(n):
    q = Queue("env_0")
    rng = np.random.default_rng(0)
    for i in range(n):
        q.push(rng.integers(0, 255, size=1000 + 37 * i, dtype=np.uint8).tobytes())


def test_ring_wraps_and_keeps_order_across_processes():
    with QueueServer(["env_0"], capacity=16 << 10):  # small ring: forces wrap-around and back-pressure
        ctx = mp.get_context("spawn")
        n = 60
        p = ctx.Process(target=_producer, args=(n,))
        p.start()
        q = Queue("env_0")
        rng = np.random.default_rng(0)
        for i in range(n):
            got = q.pop_wait(20.0)
            want = rng.integers(0, 255, size=1000 + 37 * i, dtype=np.uint8).tobytes()
            assert got == want, i
        assert q.pop() is None
        p.join(10)
        assert p.exitcode == 0
        q.close()


def test_episode_and_weights_formats_round_trip_bit_exactly():
    rng = np.random.default_rng(1)
    S, A, T = 24, 6, 17
    episode = [[rng.standard_normal(S).astype(np.float32), rng.uniform(-1, 1, A).astype(np.float32),
                float(rng.uniform()), bool(i == T - 1), rng.standard_normal(S).astype(np.float32)] for i in range(T)]
    rows, s_dim, a_dim = unpack_episode(pack_episode(episode, S, A))
    assert (s_dim, a_dim) == (S, A) and rows.shape == (T, 2 * S + A + 2)
    back = episode_rows_to_list(rows, S, A)
    for (s, a, r, d, s2), (bs, ba, br, bd, bs2) in zip(episode, back):
        assert np.array_equal(s, bs) and np.array_equal(a, ba) and np.array_equal(s2, bs2)
        assert np.float32(r) == np.float32(br) and d == bd
    sd = {"mlp.nn.0.weight": torch.randn(256, 24), "mlp.nn.0.bias": torch.randn(256), "scalar": torch.tensor(3.5)}
    out = unpack_weights(pack_weights(sd))
    assert list(out) == list(sd)
    for k in sd:
        assert torch.equal(out[k], sd[k]) and out[k].shape == sd[k].shape


def test_queue_needs_a_session_and_refuses_remote_hosts():
    import os

    import pytest

    os.environ.pop("OPRL_B200_SESSION", None)
    with pytest.raises(RuntimeError):
        Queue("env_0")
    with QueueServer(["env_0"]):
        with pytest.raises(ValueError):
            Queue("env_0", host="10.0.0.1")
