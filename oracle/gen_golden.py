"""Generate tests/golden/*.npz by running the REFERENCE implementation (imported from
/root/reference/src -- only present in the build container) on fixed seeds, minibatches
and noise, and assert that oracle/oprl_oracle.py reproduces it.

    python oracle/gen_golden.py            # writes tests/golden/{ddpg,td3,sac,sac_fixed,tqc,buffer}.npz

TEST INFRASTRUCTURE.  The fixtures it writes are what pins the oracle (and, through the
`-m gpu` tests, the CUDA engine) to the reference; the reference's own tests hold no
numerical vectors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import oprl_oracle as O  # noqa: E402


class NullLogger:
    log_dir = "/tmp"

    def log_scalar(self, *a, **k):
        pass

    def log_scalars(self, *a, **k):
        pass


def flat(module_or_params):
    ps = list(module_or_params.parameters()) if hasattr(module_or_params, "parameters") else module_or_params
    return torch.cat([p.detach().reshape(-1) for p in ps]).numpy().copy()


def make_batches(K, B, S, A, seed, int_done=False):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(K):
        s = rng.standard_normal((B, S), dtype=np.float32)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        r = rng.uniform(0, 1, (B, 1)).astype(np.float32)
        d = (rng.uniform(0, 1, (B, 1)) < 0.05).astype(np.float32)
        s2 = rng.standard_normal((B, S), dtype=np.float32)
        out.append((s, a, r, d, s2))
    return out


def ref_algo(name, S, A, **kw):
    from oprl.algos.ddpg import DDPG
    from oprl.algos.sac import SAC
    from oprl.algos.td3 import TD3
    from oprl.algos.tqc import TQC

    cls = dict(ddpg=DDPG, td3=TD3, sac=SAC, tqc=TQC)[name]
    return cls(logger=NullLogger(), state_dim=S, action_dim=A, **kw).create()


def critic_nets(name, critic):
    if name == "ddpg":
        return [critic.q1]
    if name in ("td3", "sac"):
        return [critic.q1, critic.q2]
    return list(critic.nets)


def n_noise(name):
    return dict(ddpg=0, td3=1, sac=2, tqc=2)[name]


MIN_PREACT = 5e-7  # a few times the rounding noise of an fp32-accurate GEMM on these values


def first_update_conditioning(name, S, A, B, seed, data_seed, synthetic_init, **kw):
    """min |hidden ReLU pre-activation| over every forward pass of the FIRST update."""
    torch.manual_seed(seed)
    ref = ref_algo(name, S, A, **kw)
    spec = O.AlgoSpec(algo=name, state_dim=S, action_dim=A, tune_alpha=kw.get("tune_alpha", False),
                      lr_alpha=3e-4 if name == "tqc" else 1e-3)
    if synthetic_init:
        actor0, critics0 = O.init_params(spec, seed)
    else:
        actor0 = [p.detach().clone() for p in ref.actor.parameters()]
        critics0 = [[p.detach().clone() for p in net.parameters()] for net in critic_nets(name, ref.critic)]
    orc = O.OracleAlgo(spec, actor0, critics0)
    (batch,) = make_batches(1, B, S, A, data_seed)
    torch.manual_seed(1000)
    noise = [torch.randn(B, A) for _ in range(n_noise(name))]
    O.PREACT_PROBE.update(enabled=True, min_abs=float("inf"))
    orc.update(*[torch.from_numpy(x) for x in batch], noise=noise)
    O.PREACT_PROBE["enabled"] = False
    return O.PREACT_PROBE["min_abs"]


def run_case(name, S, A, B, K, seed, subsample=1, synthetic_init=False, min_preact=MIN_PREACT, first_full=False, **kw):
    """Well-conditioned fixture.  "Parameters after one Adam step" is discontinuous where a hidden
    ReLU pre-activation sits within rounding noise of zero (the unit's gradient switches on/off and
    Adam's first step is lr * sign(g)): with ~1e6 pre-activations per update such knife edges are
    common, and no implementation that is not bit-identical to the reference's BLAS can land on the
    same side.  So the data seed is advanced until the FIRST update (the one held to the strict
    1e-5 bar) keeps every pre-activation at least MIN_PREACT away from zero; the margin actually
    seen over all K updates is recorded for the looser multi-update check."""
    for data_seed in range(seed + 1, seed + 4000, 10):
        m0 = first_update_conditioning(name, S, A, B, seed, data_seed, synthetic_init, **kw)
        if m0 < min_preact:
            continue
        O.PREACT_PROBE.update(enabled=True, min_abs=float("inf"))
        fx = _run_case(name, S, A, B, K, seed, data_seed, subsample, synthetic_init, first_full=first_full, **kw)
        O.PREACT_PROBE["enabled"] = False
        fx["min_abs_preactivation_first"] = np.float64(m0)
        fx["min_abs_preactivation_all"] = np.float64(O.PREACT_PROBE["min_abs"])
        print(f"      data seed {data_seed}: min |hidden pre-activation| first update {m0:.3e}, "
              f"all {K} updates {O.PREACT_PROBE['min_abs']:.3e}")
        return fx
    raise RuntimeError("no well-conditioned data seed found")


def _run_case(name, S, A, B, K, seed, data_seed, subsample=1, synthetic_init=False, first_full=False, **kw):
    """Returns the fixture dict.  K updates; dumps after update 1 and after update K."""
    torch.manual_seed(seed)
    ref = ref_algo(name, S, A, **kw)
    spec = O.AlgoSpec(algo=name, state_dim=S, action_dim=A,
                      tune_alpha=kw.get("tune_alpha", False),
                      lr_alpha=3e-4 if name == "tqc" else 1e-3)
    if synthetic_init:
        actor0, critics0 = O.init_params(spec, seed)
        with torch.no_grad():
            for p, q in zip(ref.actor.parameters(), actor0):
                p.copy_(q)
            for net, qs in zip(critic_nets(name, ref.critic), critics0):
                for p, q in zip(net.parameters(), qs):
                    p.copy_(q)
            for net, qs in zip(critic_nets(name, ref.critic_target), critics0):
                for p, q in zip(net.parameters(), qs):
                    p.copy_(q)
    actor0 = [p.detach().clone() for p in ref.actor.parameters()]
    critics0 = [[p.detach().clone() for p in net.parameters()] for net in critic_nets(name, ref.critic)]
    orc = O.OracleAlgo(spec, actor0, critics0)

    fx = dict(algo=name, S=S, A=A, B=B, K=K, seed=seed, data_seed=data_seed, subsample=subsample,
              synthetic_init=int(synthetic_init), tune_alpha=int(spec.tune_alpha),
              torch_version=torch.__version__)
    if not synthetic_init:
        fx["actor0"] = flat(actor0)
        fx["critic0"] = flat([p for net in critics0 for p in net])

    batches = make_batches(K, B, S, A, data_seed)
    max_dev = 0.0
    for k, (s, a, r, d, s2) in enumerate(batches):
        # the reference draws its noise from torch's global generator; pre-draw the same stream
        torch.manual_seed(1000 + k)
        noise = [torch.randn(B, A) for _ in range(n_noise(name))]
        torch.manual_seed(1000 + k)
        ts = [torch.from_numpy(x) for x in (s, a, r, d, s2)]
        ref.update(*ts)
        sc = orc.update(*ts, noise=noise)
        for i, nz in enumerate(noise):
            fx[f"noise{k}_{i}"] = nz.numpy()
        for nm, x in zip(("s", "a", "r", "d", "s2"), (s, a, r, d, s2)):
            fx[f"{nm}{k}"] = x
        for key, val in sc.items():
            fx[f"scalar{k}_{key}"] = np.float64(val)
        if k in (0, K - 1):
            tag = "first" if k == 0 else "last"
            got = dict(
                actor=flat(ref.actor),
                critic=flat(ref.critic),
                critic_target=flat(ref.critic_target),
            )
            if spec.has_actor_target:
                got["actor_target"] = flat(ref.actor_target)
            for key, val in got.items():
                dev = float(np.abs(val - orc.flat(key)).max())
                max_dev = max(max_dev, dev)
                fx[f"{tag}_{key}_l2"] = np.float64(np.sqrt((val.astype(np.float64) ** 2).sum()))
                fx[f"{tag}_{key}_sum"] = np.float64(val.astype(np.float64).sum())
                fx[f"{tag}_{key}"] = val[::subsample].copy()
                # first_full: the COMPLETE parameter vectors after the first update as well, so that the GPU parity
                # test measures the true L2 instead of scaling a subsample (the target nets are derived in the test:
                # Polyak of the seeded init and these)
                if first_full and k == 0 and key in ("actor", "critic"):
                    fx[f"firstfull_{key}"] = val.copy()
            if k == 0:
                # gradients of the first update (reference leaves them in .grad)
                ga = torch.cat([p.grad.reshape(-1) for p in ref.actor.parameters()]).numpy()
                dev = float(np.abs(ga - np.concatenate([g.reshape(-1).numpy() for g in orc.last_actor_grads])).max())
                max_dev = max(max_dev, dev)
                fx["first_actor_grad"] = ga[::subsample].copy()
                # NB: critic .grad after update() holds the actor-step leftovers (ddpg.py:104-106),
                # so the critic-step gradient comes from the (just validated) oracle.
                fx["first_critic_grad"] = np.concatenate(
                    [g.reshape(-1).numpy() for g in orc.last_critic_grads])[::subsample].copy()
            if spec.tune_alpha:
                fx[f"{tag}_log_alpha"] = np.float64(ref.log_alpha.item())
                dev = abs(ref.log_alpha.item() - orc.log_alpha.item())
                max_dev = max(max_dev, dev)
    fx["oracle_vs_reference_max_abs_dev"] = np.float64(max_dev)
    print(f"{name:5s} S={S} A={A} B={B} K={K}: oracle vs reference max |dev| = {max_dev:.3e}")
    assert max_dev <= 1e-6, "oracle does not reproduce the reference"
    return fx


def buffer_case():
    """Replay path: add_episode / add_transition bookkeeping + sample() indices, from the
    reference EpisodicReplayBuffer with ragged episodes and ring wrap-around."""
    from oprl.buffers.episodic_buffer import EpisodicReplayBuffer

    S, A, L = 5, 3, 20
    buf = EpisodicReplayBuffer(buffer_size_transitions=8 * L, state_dim=S, action_dim=A,
                               max_episode_lenth=L).create()
    for k in buf._tensors:
        buf._tensors[k].zero_()
    rng = np.random.default_rng(7)
    log = []
    lens = [20, 7, 13, 20, 1, 9, 20, 16, 11, 20, 5]  # wraps the 8-episode ring
    for ep_i, n in enumerate(lens):
        for t_ in range(n):
            s = rng.standard_normal(S).astype(np.float32)
            a = rng.uniform(-1, 1, A).astype(np.float32)
            r = float(rng.uniform())
            d = False
            buf.add_transition(s, a, r, d, episode_done=(t_ == n - 1))
            log.append((s, a, r, d, t_ == n - 1))
    fx = dict(S=S, A=A, L=L, cap=8 * L, lens=np.array(lens))
    fx["log_s"] = np.stack([x[0] for x in log])
    fx["log_a"] = np.stack([x[1] for x in log])
    fx["log_r"] = np.array([x[2] for x in log], np.float64)
    fx["log_d"] = np.array([x[3] for x in log], np.bool_)
    fx["log_ep_done"] = np.array([x[4] for x in log], np.bool_)
    fx["ep_lens"] = np.array(buf.ep_lens)
    fx["ep_pointer"] = np.int64(buf._ep_pointer)
    fx["episodes_counter"] = np.int64(buf.episodes_counter)
    fx["n_transitions"] = np.int64(len(buf))
    fx["states"] = buf.states.numpy().copy()
    fx["actions"] = buf.actions.numpy().copy()
    fx["rewards"] = buf.rewards.numpy().copy()
    fx["dones"] = buf.dones.numpy().copy()
    np.random.seed(0)
    inds = np.random.randint(low=0, high=len(buf), size=64)
    np.random.seed(0)
    out = buf.sample(64)
    ep, step = buf._inds_to_episodic(inds)
    oep, ostep = O.inds_to_episodic(inds, buf.ep_lens, buf.episodes_counter)
    assert (ep == oep).all() and (step == ostep).all()
    fx["inds"], fx["ep"], fx["step"] = inds, ep, step
    for nm, x in zip(("s", "a", "r", "d", "s2"), out):
        fx["batch_" + nm] = x.numpy().copy()
    og = O.gather_batch(fx["states"], fx["actions"], fx["rewards"], fx["dones"], oep, ostep)
    for x, y in zip(out, og):
        assert (x.numpy() == y).all()
    print("buffer: oracle index math + gather bit-exact vs reference")
    return fx


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(1)  # reference update is bit-identical across thread counts given init
    cases = {
        # BASELINE.json configs[0]: DDPG walker-walk (S=24, A=6), batch 256
        "ddpg": lambda: run_case("ddpg", 24, 6, 256, 5, 0),
        # configs[1]: TD3 cheetah-run (S=17, A=6), batch 256; K=4 covers delayed actor twice
        "td3": lambda: run_case("td3", 17, 6, 256, 4, 1),
        # configs[2]: SAC humanoid-stand (S=67, A=21), auto-alpha, batch 1024
        "sac": lambda: run_case("sac", 67, 21, 1024, 3, 2, tune_alpha=True),
        # fixed alpha + the reference test's batch of 8 (tests/functional/test_rl_algos.py:24)
        "sac_fixed": lambda: run_case("sac", 24, 6, 8, 3, 3, subsample=8, tune_alpha=False),
        # configs[3]: TQC walker-walk 5x25, batch 256; 2.8M params -> seeded init + 1/64 subsample
        # (5.9M pre-activations per update: the margin that can be found is smaller)
        "tqc": lambda: run_case("tqc", 24, 6, 256, 2, 4, subsample=64, synthetic_init=True, min_preact=1.2e-7, first_full=True),
        # ragged batch of 8 on DDPG (reference test shape)
        "ddpg_b8": lambda: run_case("ddpg", 24, 6, 8, 2, 5, subsample=8),
    }
    only = sys.argv[1:]
    for name, fn in cases.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **fn())
    if not only or "buffer" in only:
        np.savez_compressed(os.path.join(out_dir, "buffer.npz"), **buffer_case())


if __name__ == "__main__":
    main()
