"""CPU: the oracle (oracle/oprl_oracle.py) against the fixtures generated from the reference
implementation itself (oracle/gen_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import oprl_oracle as O
from tests.util import load_case, oracle_from_fixture, run_fixture_updates

CASES = ["ddpg", "ddpg_b8", "td3", "sac", "sac_fixed", "tqc"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_fixture(name):
    torch.set_num_threads(1)
    fx = load_case(name)
    assert float(fx["oracle_vs_reference_max_abs_dev"]) <= 1e-6  # recorded at generation time
    orc = oracle_from_fixture(fx)
    sub = int(fx["subsample"])
    K = int(fx["K"])

    def check(tag):
        for key in ("actor", "critic", "critic_target", "actor_target"):
            k = f"{tag}_{key}"
            if k not in fx:
                continue
            got = orc.flat(key)
            np.testing.assert_allclose(got[::sub], fx[k], rtol=0, atol=2e-7, err_msg=k)
            l2 = np.sqrt((got.astype(np.float64) ** 2).sum())
            assert abs(l2 - float(fx[k + "_l2"])) <= 1e-6 * max(1.0, l2)
        if f"{tag}_log_alpha" in fx:
            assert abs(orc.log_alpha.item() - float(fx[f"{tag}_log_alpha"])) <= 1e-9

    for k in range(K):
        sc = run_fixture_updates(orc, fx, k)
        for key, val in sc.items():
            ref = float(fx[f"scalar{k}_{key}"])
            assert abs(val - ref) <= 1e-6 * max(1.0, abs(ref)), (k, key)
        if k == 0:
            check("first")
            ga = np.concatenate([g.reshape(-1).numpy() for g in orc.last_actor_grads])
            np.testing.assert_allclose(ga[::sub], fx["first_actor_grad"], rtol=0, atol=1e-8)
    check("last")


def test_buffer_index_math_and_gather_bit_exact():
    fx = load_case("buffer")
    ep, step = O.inds_to_episodic(fx["inds"], list(fx["ep_lens"]), int(fx["episodes_counter"]))
    assert (ep == fx["ep"]).all() and (step == fx["step"]).all()
    out = O.gather_batch(fx["states"], fx["actions"], fx["rewards"], fx["dones"], ep, step)
    for nm, x in zip(("s", "a", "r", "d", "s2"), out):
        assert (x == fx["batch_" + nm]).all()


def test_oracle_header_declares_test_infrastructure():
    src = open(os.path.join(os.path.dirname(O.__file__), "oprl_oracle.py")).read()
    assert "TEST INFRASTRUCTURE ONLY" in src


def test_nstep_oracle_reduces_to_the_reference_batch_and_telescopes():
    """The n-step extension of the gather (oracle.nstep_batch): n = 1 is the reference's batch bit for bit, and the
    1-step target on the assembled batch equals the n-step target (constant bootstrap value)."""
    import numpy as np

    from oracle import oprl_oracle as O

    rng = np.random.default_rng(0)
    E, Lmax, S, A = 6, 12, 5, 2
    st = rng.standard_normal((E, Lmax + 1, S)).astype(np.float32)
    ac = rng.uniform(-1, 1, (E, Lmax, A)).astype(np.float32)
    rw = rng.uniform(0, 1, (E, Lmax, 1)).astype(np.float32)
    dn = (rng.uniform(0, 1, (E, Lmax, 1)) < 0.15).astype(np.float32)
    ep_lens = [12, 7, 12, 3, 9, 12]
    ep = np.array([0, 1, 1, 3, 4, 5, 2, 0])
    step = np.array([0, 5, 6, 2, 8, 10, 4, 11])
    s, a, r, d, s2 = O.nstep_batch(st, ac, rw, dn, ep_lens, ep, step, 1, 0.99)
    assert np.array_equal(s, st[ep, step]) and np.array_equal(a, ac[ep, step]) and np.array_equal(r, rw[ep, step])
    assert np.array_equal(d, dn[ep, step]) and np.array_equal(s2, st[ep, step + 1])
    gamma, q = 0.9, 3.0  # constant bootstrap value: compare against the explicit n-step sum in float64
    for n in (2, 3, 5):
        s, a, r, d, s2 = O.nstep_batch(st, ac, rw, dn, ep_lens, ep, step, n, gamma)
        for i in range(len(ep)):
            e, t0 = int(ep[i]), int(step[i])
            m, R, done = 0, 0.0, 0.0
            while m < n and t0 + m < ep_lens[e]:
                R += gamma ** m * float(rw[e, t0 + m, 0])
                done = float(dn[e, t0 + m, 0])
                m += 1
                if done:
                    break
            want = R + (1.0 - done) * gamma ** m * q
            got = float(r[i, 0]) + (1.0 - float(d[i, 0])) * gamma * q
            assert abs(got - want) < 1e-5, (n, i, got, want)
            assert np.array_equal(s2[i], st[e, t0 + m])
