"""GPU parity: the CUDA engine (through the C ABI / python shim) against the fixtures produced
by the reference itself, and against the CPU oracle on fresh seeded inputs.

Bars (BASELINE.json north_star): |loss - loss_ref| <= 1e-4 and ||theta - theta_ref||_2 <= 1e-5
over the concatenation of all networks after one update, fixed seeds + fixed minibatch."""
import numpy as np
import pytest
import torch

from tests.util import (fixture_batch, fixture_noise, load_case, oracle_from_fixture,
                        run_fixture_updates, spec_from_fixture)

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-4
PARAM_L2_TOL = 1e-5


def make_algo(fx, **kw):
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.sac import SAC
    from oprl_b200.algos.td3 import TD3
    from oprl_b200.algos.tqc import TQC

    spec = spec_from_fixture(fx)

    class NullLogger:
        log_dir = "/tmp"

        def log_scalar(self, *a, **k):
            pass

        def log_scalars(self, *a, **k):
            pass

    cls = dict(ddpg=DDPG, td3=TD3, sac=SAC, tqc=TQC)[spec.algo]
    if spec.algo == "sac":
        kw.setdefault("tune_alpha", spec.tune_alpha)
    algo = cls(logger=NullLogger(), state_dim=spec.state_dim, action_dim=spec.action_dim, **kw).create()
    return algo


def load_initial(algo, orc):
    ar = algo.engine.arena
    for name in ("actor", "critic"):
        flat = torch.from_numpy(orc.flat(name)).cuda()
        ar[name]["theta"].copy_(flat)
        if ar[name]["target"] is not None:
            ar[name]["target"].copy_(flat)
    algo.engine.mark_params_dirty()


def engine_flat(algo, which):
    ar = algo.engine.arena
    name, key = {"actor": ("actor", "theta"), "critic": ("critic", "theta"),
                 "actor_target": ("actor", "target"), "critic_target": ("critic", "target")}[which]
    t_ = ar[name][key]
    return None if t_ is None else t_.detach().cpu().numpy()


def compare_to_fixture(algo, fx, tag):
    sub = int(fx["subsample"])
    sq = 0.0
    for key in ("actor", "critic", "critic_target", "actor_target"):
        k = f"{tag}_{key}"
        if k not in fx:
            continue
        got = engine_flat(algo, key)[::sub]
        sq += float(((got.astype(np.float64) - fx[k].astype(np.float64)) ** 2).sum())
    return np.sqrt(sq * sub)  # subsampled fixtures: scale to the full-vector estimate


@pytest.mark.parametrize("name", ["ddpg", "ddpg_b8", "td3", "sac", "sac_fixed", "tqc"])
def test_fixture_parity(name):
    fx = load_case(name)
    orc = oracle_from_fixture(fx)
    algo = make_algo(fx)
    load_initial(algo, orc)
    K = int(fx["K"])
    for k in range(K):
        noise = fixture_noise(fx, k)
        for i, nz in enumerate(noise):
            algo.engine.set_noise(i, nz)
        batch = [x.cuda() for x in fixture_batch(fx, k)]
        algo.update(*batch)
        sc = algo.engine.scalars()
        for key in ("critic_loss", "actor_loss", "alpha_loss", "alpha"):
            fk = f"scalar{k}_{key}"
            if fk in fx:
                assert abs(sc[key] - float(fx[fk])) <= LOSS_TOL, (k, key, sc[key], float(fx[fk]))
        if k == 0:
            l2 = compare_to_fixture(algo, fx, "first")
            print(f"{name}: param L2 after 1 update = {l2:.3e}")
            assert l2 <= PARAM_L2_TOL
            if "first_log_alpha" in fx:
                assert abs(algo.engine.state().log_alpha - float(fx["first_log_alpha"])) <= 1e-7
    l2 = compare_to_fixture(algo, fx, "last")
    margin = float(fx["min_abs_preactivation_all"])
    print(f"{name}: param L2 after {K} updates = {l2:.3e} (min |pre-activation| over them {margin:.1e})")
    # Later updates are not conditioned (oracle/gen_golden.py): a hidden pre-activation within
    # rounding noise of zero flips a ReLU and with it the sign of a few near-zero Adam steps
    # (2*lr = 6e-4 each).  Strict bar when the fixture stayed clear of that, loose bound otherwise.
    assert l2 <= (PARAM_L2_TOL * K if margin >= 5e-7 else 5e-3)


def test_host_batch_and_int64_done_match_the_device_path():
    """update() takes what the reference's own test feeds it (tests/functional/test_rl_algos.py:25-31):
    CPU tensors, an int64 `done`, next_state aliasing state -- through the pinned staging ring and the
    copy-stream load into the other working set."""
    fx = load_case("ddpg_b8")
    results = []
    for on_host in (False, True):
        orc = oracle_from_fixture(fx)
        algo = make_algo(fx)
        load_initial(algo, orc)
        s, a, r, d, _ = fixture_batch(fx, 0)
        d = d.to(torch.int64)
        batch = [s, a, r, d, s] if on_host else [x.cuda() for x in (s, a, r, d, s)]
        for _ in range(3):
            algo.update(*batch)
        torch.cuda.synchronize()
        results.append({g: ar["theta"].clone() for g, ar in algo.engine.arena.items()})
    for g in results[0]:
        assert torch.equal(results[0][g], results[1][g]), g


@pytest.mark.parametrize("name,B", [("ddpg", 1), ("ddpg", 100), ("td3", 257), ("sac", 33)])
def test_odd_batch_sizes_against_the_oracle(name, B):
    """Batch sizes that are not multiples of the 128-row MMA tile (or of anything): fresh seeded
    inputs, the CPU oracle as the checker.  The data seed is advanced until the update is well
    conditioned (no ReLU pre-activation within 5e-7 of zero, see oracle/gen_golden.py)."""
    from oracle import oprl_oracle as O

    fx = load_case({"ddpg": "ddpg_b8", "td3": "td3", "sac": "sac_fixed"}[name])
    spec = spec_from_fixture(fx)
    S, A = spec.state_dim, spec.action_dim
    n_noise = {"ddpg": 0, "td3": 1, "sac": 2}[name]
    for seed in range(100, 160):
        rng = np.random.default_rng(seed)
        batch = [torch.from_numpy(x) for x in (
            rng.standard_normal((B, S), dtype=np.float32), rng.uniform(-1, 1, (B, A)).astype(np.float32),
            rng.uniform(0, 1, (B, 1)).astype(np.float32), (rng.uniform(0, 1, (B, 1)) < 0.3).astype(np.float32),
            rng.standard_normal((B, S), dtype=np.float32))]
        g = torch.Generator().manual_seed(seed)
        noise = [torch.randn(B, A, generator=g) for _ in range(n_noise)]
        orc = oracle_from_fixture(fx)
        O.PREACT_PROBE.update(enabled=True, min_abs=float("inf"))
        ref = orc.update(*batch, noise=noise)
        O.PREACT_PROBE["enabled"] = False
        if O.PREACT_PROBE["min_abs"] >= 5e-7:
            break
    else:
        pytest.skip("no well-conditioned seed found")
    algo = make_algo(fx)
    load_initial(algo, oracle_from_fixture(fx))
    for i, nz in enumerate(noise):
        algo.engine.set_noise(i, nz)
    algo.update(*[x.cuda() for x in batch])
    sc = algo.engine.scalars()
    for key in ("critic_loss", "actor_loss"):
        assert abs(sc[key] - ref[key]) <= LOSS_TOL, (key, sc[key], ref[key])
    sq = 0.0
    for key in ("actor", "critic", "critic_target", "actor_target"):
        got = engine_flat(algo, key)
        if got is not None:
            sq += float(((got.astype(np.float64) - orc.flat(key)) ** 2).sum())
    print(f"{name} B={B}: param L2 after 1 update vs oracle = {np.sqrt(sq):.3e}")
    assert np.sqrt(sq) <= PARAM_L2_TOL
