"""CPU oracle for the off-policy update hot path of schatty/oprl.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oprl_b200/`` may import this module:
it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` can check (or time) the algorithm the
CUDA engine implements.  The product path fails loudly without its CUDA library.

What it is
----------
A functional restatement, on CPU PyTorch fp32 tensors, of exactly the arithmetic the
reference performs for   sample -> target-Q -> critic step -> actor step -> Polyak:

* the reference is pure Python; all of its arithmetic is the third-party dependency
  ``torch`` (pinned ``torch==2.2.2``, reference pyproject.toml:21; this image has
  2.11.0 -- same operators, same Adam formula), plus ``numpy`` (pinned 1.26.4) for
  the replay index math.  So the restatement keeps the same operator sequence
  (``addmm`` layers, autograd backward, single-tensor Adam) and is therefore also a
  fair stand-in for the reference's CPU cost (bench.py ``cpu_baseline`` kind="port").
* every function cites the reference file:line it follows (paths relative to the
  reference root, ``src/oprl/...``).

Parity pinning
--------------
The reference's own tests hold no numerical fixtures, so the
oracle is pinned against outputs of the reference itself: ``oracle/gen_golden.py``
imports the reference from /root/reference (build container only), runs it on fixed
seeds / minibatches / noise, asserts this oracle reproduces it, and commits the
vectors under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle
against those vectors on every run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0  # nn_models.py:11


# --------------------------------------------------------------------------- specs
@dataclass
class AlgoSpec:
    """Hyper-parameters; defaults are the reference dataclass defaults."""

    algo: str  # "ddpg" | "td3" | "sac" | "tqc"
    state_dim: int
    action_dim: int
    gamma: float = 0.99
    tau: float = 5e-3
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    lr_alpha: float = 1e-3  # sac.py:25 ; tqc.py:69 uses 3e-4
    policy_noise: float = 0.2  # td3.py:21
    noise_clip: float = 0.5  # td3.py:23
    policy_freq: int = 2  # td3.py:24
    max_action: float = 1.0
    tune_alpha: bool = False  # sac.py:22 (TQC: always on)
    alpha_init: float = 0.2  # sac.py:27 ; tqc.py:105
    n_quantiles: int = 25  # tqc.py:72
    n_nets: int = 5  # tqc.py:73
    top_quantiles_to_drop: int = 2  # tqc.py:71
    actor_hidden: tuple = (256, 256)
    critic_hidden: tuple = (256, 256)  # TQC: (512, 512, 512), tqc.py:49

    def __post_init__(self):
        if self.algo == "tqc":
            self.tune_alpha = True
            if self.critic_hidden == (256, 256):
                self.critic_hidden = (512, 512, 512)

    # layer dims ------------------------------------------------------------
    @property
    def n_critics(self) -> int:
        return {"ddpg": 1, "td3": 2, "sac": 2, "tqc": self.n_nets}[self.algo]

    @property
    def actor_out(self) -> int:
        return self.action_dim if self.algo in ("ddpg", "td3") else 2 * self.action_dim

    @property
    def critic_out(self) -> int:
        return self.n_quantiles if self.algo == "tqc" else 1

    def actor_dims(self):
        return [self.state_dim, *self.actor_hidden, self.actor_out]

    def critic_dims(self):
        return [self.state_dim + self.action_dim, *self.critic_hidden, self.critic_out]

    @property
    def has_actor_target(self) -> bool:
        return self.algo in ("ddpg", "td3")

    @property
    def target_entropy(self) -> float:
        return -float(self.action_dim)  # sac.py:70, tqc.py:90


def mlp_param_shapes(dims):
    """Parameter order of nn_models.MLP.parameters(): per layer weight [out,in], bias [out]
    (nn_models.py:98-104)."""
    shapes = []
    for i in range(len(dims) - 1):
        shapes.append((dims[i + 1], dims[i]))
        shapes.append((dims[i + 1],))
    return shapes


# ------------------------------------------------------------------------ networks
# Conditioning probe (gen_golden.py): smallest |pre-activation| seen by any hidden ReLU while
# enabled.  A pre-activation within fp32 rounding noise of 0 makes "parameters after one Adam
# step" ill-conditioned (the unit's gradient switches on/off and Adam's first step is
# lr * sign(g)), so fixtures are generated from seeds that stay clear of that knife edge.
PREACT_PROBE = {"enabled": False, "min_abs": float("inf")}


def mlp_forward(params, x):
    """nn_models.MLP.forward (nn_models.py:106-107): Linear -> ReLU ... -> Linear."""
    n_layers = len(params) // 2
    for i in range(n_layers):
        x = torch.addmm(params[2 * i + 1], x, params[2 * i].t())
        if i < n_layers - 1:
            if PREACT_PROBE["enabled"]:
                ax = x.detach().abs()
                PREACT_PROBE["min_abs"] = min(PREACT_PROBE["min_abs"], float(ax.min()))
                # exact zeros are structural (a row whose inputs and bias are all zero), not a rounding knife edge
                nz = ax[ax > 0]
                if nz.numel():
                    PREACT_PROBE["min_abs_nonzero"] = min(PREACT_PROBE.get("min_abs_nonzero", float("inf")), float(nz.min()))
            x = torch.relu(x)
    return x


def critic_forward(nets, s, a):
    """Critic / DoubleCritic / QuantileQritic forward (nn_models.py:43-45,74-77, tqc.py:55-58):
    every net sees cat([s, a], -1)."""
    x = torch.cat([s, a], dim=-1)
    return [mlp_forward(p, x) for p in nets]


def deterministic_policy(params, s):
    """DeterministicPolicy.forward (nn_models.py:135-136)."""
    return torch.tanh(mlp_forward(params, s))


def gaussian_policy(params, s, eps, action_dim):
    """GaussianActor.forward in training mode + TanhNormal (nn_models.py:169-182,197-214).
    ``eps`` is the standard-normal draw the reference takes from the global generator."""
    out = mlp_forward(params, s)
    mean, log_std = out[:, :action_dim], out[:, action_dim:]
    log_std = log_std.clamp(LOG_STD_MIN, LOG_STD_MAX)
    std = torch.exp(log_std)
    pre = mean + std * eps
    action = torch.tanh(pre)
    # TanhNormal.log_prob (nn_models.py:208-210): log_det is built first, then
    # Normal.log_prob = -((x-mu)^2)/(2 var) - log(std) - log(sqrt(2 pi)); keeping the reference's
    # node-creation order keeps autograd's multi-path accumulation order (bit-exact grads).
    log_det = (
        2 * np.log(2)
        + torch.nn.functional.logsigmoid(2 * pre)
        + torch.nn.functional.logsigmoid(-2 * pre)
    )
    var = std**2
    log_scale = std.log()
    normal_lp = -((pre - mean) ** 2) / (2 * var) - log_scale - math.log(math.sqrt(2 * math.pi))
    log_prob = (normal_lp - log_det).sum(dim=1, keepdim=True)
    return action, log_prob


def quantile_huber_loss(quantiles, samples):
    """tqc.py:14-36.  quantiles [B, nets, nq]; samples [B, kept]."""
    delta = samples[:, None, None, :] - quantiles[:, :, :, None]
    abs_delta = torch.abs(delta)
    huber = torch.where(abs_delta > 1, abs_delta - 0.5, delta**2 * 0.5)
    nq = quantiles.shape[2]
    tau = torch.arange(nq).float() / nq + 1 / 2 / nq
    return (torch.abs(tau[None, None, :, None] - (delta < 0).float()) * huber).mean()


# ---------------------------------------------------------------------------- Adam
@dataclass
class AdamState:
    lr: float
    m: list
    v: list
    step: int = 0
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8


def adam_init(params, lr):
    return AdamState(lr, [torch.zeros_like(p) for p in params], [torch.zeros_like(p) for p in params])


def adam_step(params, grads, st: AdamState):
    """torch.optim.Adam, single-tensor path, defaults (constructed at ddpg.py:51,56;
    torch 2.2.2 torch/optim/adam.py::_single_tensor_adam):
        m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, value=1-b2)
        denom = (v.sqrt() / sqrt(1-b2^t)).add_(eps); p.addcdiv_(m, denom, value=-lr/(1-b1^t))
    """
    st.step += 1
    bc1 = 1 - st.beta1**st.step
    bc2 = 1 - st.beta2**st.step
    step_size = st.lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    with torch.no_grad():
        for p, g, m, v in zip(params, grads, st.m, st.v):
            m.lerp_(g, 1 - st.beta1)
            v.mul_(st.beta2).addcmul_(g, g, value=1 - st.beta2)
            denom = (v.sqrt() / bc2_sqrt).add_(st.eps)
            p.addcdiv_(m, denom, value=-step_size)


def polyak(target, source, tau):
    """ddpg.py:72-84 / tqc.py:154-159 (tau*p + (1-tau)*t) and nn_functions.soft_update
    (nn_functions.py:5-10: t.mul_(1-tau); t.add_(tau*p)) -- the same three fp32 roundings."""
    with torch.no_grad():
        for t_, p in zip(target, source):
            t_.copy_(tau * p + (1 - tau) * t_)


# -------------------------------------------------------------------- the algorithm
def _leaf(p):
    return p.detach().clone().requires_grad_(True)


class OracleAlgo:
    """State + update() of one algorithm.  Parameters are flat lists in ``parameters()`` order."""

    def __init__(self, spec: AlgoSpec, actor, critics):
        self.spec = spec
        self.actor = [_leaf(p) for p in actor]
        self.critics = [[_leaf(p) for p in net] for net in critics]
        self.actor_target = [p.detach().clone() for p in actor] if spec.has_actor_target else None
        self.critics_target = [[p.detach().clone() for p in net] for net in critics]
        self.opt_actor = adam_init(self.actor, spec.lr_actor)
        self.opt_critic = adam_init(self._critic_flat(), spec.lr_critic)
        self.alpha = float(spec.alpha_init)
        self.log_alpha = None
        if spec.tune_alpha:
            # float64 0-dim tensor, as sac.py:66-68 / tqc.py:105
            self.log_alpha = torch.tensor(np.log(spec.alpha_init), requires_grad=True)
            self.opt_alpha = adam_init([self.log_alpha], spec.lr_alpha)
        self.update_step = 0
        self.scalars = {}

    def _critic_flat(self):
        return [p for net in self.critics for p in net]

    def _critic_target_flat(self):
        return [p for net in self.critics_target for p in net]

    # ------------------------------------------------------------------ dispatch
    def update(self, s, a, r, d, s2, noise=()):
        """AlgorithmProtocol.update (algos/protocols.py:31-38).  ``noise``: the standard-normal
        draws the reference would take from torch's global generator, in call order."""
        d = d.to(torch.float32) if d.dtype != torch.float32 else d
        return getattr(self, "_update_" + self.spec.algo)(s, a, r, d, s2, list(noise))

    def _critic_step(self, loss):
        flat = self._critic_flat()
        grads = torch.autograd.grad(loss, flat)
        self.last_critic_grads = [g.clone() for g in grads]
        adam_step(flat, grads, self.opt_critic)

    def _actor_step(self, loss):
        grads = torch.autograd.grad(loss, self.actor)
        self.last_actor_grads = [g.clone() for g in grads]
        adam_step(self.actor, grads, self.opt_actor)

    # ---------------------------------------------------------------------- DDPG
    def _update_ddpg(self, s, a, r, d, s2, noise):
        sp = self.spec
        # _update_critic, ddpg.py:86-101
        with torch.no_grad():
            a2 = deterministic_policy(self.actor_target, s2)
            tq = critic_forward(self.critics_target, s2, a2)[0]
            y = r + (1.0 - d) * sp.gamma * tq
        q = critic_forward(self.critics, s, a)[0]
        critic_loss = (q - y).pow(2).mean()
        self._critic_step(critic_loss)
        # _update_actor, ddpg.py:103-107
        actor_loss = -critic_forward(self.critics, s, deterministic_policy(self.actor, s))[0].mean()
        self._actor_step(actor_loss)
        # Polyak, ddpg.py:72-84 (critic, then actor)
        polyak(self._critic_target_flat(), self._critic_flat(), sp.tau)
        polyak(self.actor_target, self.actor, sp.tau)
        self.scalars = dict(critic_loss=critic_loss.item(), actor_loss=actor_loss.item(),
                            q_mean=q.mean().item(), q_target_mean=y.mean().item())
        return self.scalars

    # ----------------------------------------------------------------------- TD3
    def _update_td3(self, s, a, r, d, s2, noise):
        sp = self.spec
        # _update_critic, td3.py:87-116
        q1, q2 = critic_forward(self.critics, s, a)
        with torch.no_grad():
            n = (noise[0] * sp.policy_noise).clamp(-sp.noise_clip, sp.noise_clip)
            a2 = (deterministic_policy(self.actor_target, s2) + n).clamp(-sp.max_action, sp.max_action)
            q1n, q2n = critic_forward(self.critics_target, s2, a2)
            qn = torch.min(q1n, q2n)
        y = r + (1.0 - d) * sp.gamma * qn
        critic_loss = (q1 - y).pow(2).mean() + (q2 - y).pow(2).mean()
        self._critic_step(critic_loss)
        self.scalars = dict(critic_loss=critic_loss.item(), q_mean=q1.mean().item(),
                            q_target_mean=y.mean().item())
        # delayed actor + both Polyaks, td3.py:81-85,134-141
        if self.update_step % sp.policy_freq == 0:
            q_pi = critic_forward(self.critics[:1], s, deterministic_policy(self.actor, s))[0]
            actor_loss = -q_pi.mean()
            self._actor_step(actor_loss)
            polyak(self._critic_target_flat(), self._critic_flat(), sp.tau)
            polyak(self.actor_target, self.actor, sp.tau)
            self.scalars["actor_loss"] = actor_loss.item()
        self.update_step += 1
        return self.scalars

    # ----------------------------------------------------------------------- SAC
    def _update_sac(self, s, a, r, d, s2, noise):
        sp = self.spec
        A = sp.action_dim
        # update_critic, sac.py:88-110 (the ONLINE actor proposes the next action)
        q1, q2 = critic_forward(self.critics, s, a)
        with torch.no_grad():
            a2, logp2 = gaussian_policy(self.actor, s2, noise[0], A)
            q1n, q2n = critic_forward(self.critics_target, s2, a2)
            qn = torch.min(q1n, q2n) - self.alpha * logp2
        y = r + (1.0 - d) * sp.gamma * qn
        critic_loss = (q1 - y).pow(2).mean() + (q2 - y).pow(2).mean()
        self._critic_step(critic_loss)
        # update_actor, sac.py:123-141
        a_pi, logp = gaussian_policy(self.actor, s, noise[1], A)
        qs1, qs2 = critic_forward(self.critics, s, a_pi)
        actor_loss = self.alpha * logp.mean() - torch.min(qs1, qs2).mean()
        self._actor_step(actor_loss)
        self.scalars = dict(critic_loss=critic_loss.item(), actor_loss=actor_loss.item(),
                            q_mean=q1.mean().item(), q_target_mean=y.mean().item(),
                            logpi_mean=logp.mean().item())
        if sp.tune_alpha:
            loss_alpha = -self.log_alpha * (sp.target_entropy + logp.detach().mean())
            (g,) = torch.autograd.grad(loss_alpha, [self.log_alpha])
            adam_step([self.log_alpha], [g], self.opt_alpha)
            self.alpha = self.log_alpha.detach().exp().item()
            self.scalars["alpha_loss"] = loss_alpha.item()
        self.scalars["alpha"] = self.alpha
        # soft_update(critic_target, critic), sac.py:85
        polyak(self._critic_target_flat(), self._critic_flat(), sp.tau)
        self.update_step += 1
        return self.scalars

    # ----------------------------------------------------------------------- TQC
    def _update_tqc(self, s, a, r, d, s2, noise):
        sp = self.spec
        A = sp.action_dim
        B = s.shape[0]
        alpha = torch.exp(self.log_alpha)  # float64 0-dim, tqc.py:126
        with torch.no_grad():  # tqc.py:129-145
            a2, logp2 = gaussian_policy(self.actor, s2, noise[0], A)
            nz = torch.stack(critic_forward(self.critics_target, s2, a2), dim=1)
            sz, _ = torch.sort(nz.reshape(B, -1))
            keep = sp.n_quantiles * sp.n_nets - sp.top_quantiles_to_drop
            sz = sz[:, :keep]
            target = r + (1 - d) * sp.gamma * (sz - alpha * logp2)
        cur = torch.stack(critic_forward(self.critics, s, a), dim=1)
        critic_loss = quantile_huber_loss(cur, target)
        self._critic_step(critic_loss)
        # critic Polyak BEFORE the actor step, tqc.py:154-159
        polyak(self._critic_target_flat(), self._critic_flat(), sp.tau)
        # policy and alpha loss, tqc.py:162-177
        a_pi, logp = gaussian_policy(self.actor, s, noise[1], A)
        alpha_loss = -self.log_alpha * (logp + sp.target_entropy).detach().mean()
        z_pi = torch.stack(critic_forward(self.critics, s, a_pi), dim=1)
        actor_loss = (alpha * logp - z_pi.mean(2).mean(1, keepdim=True)).mean()
        self._actor_step(actor_loss)
        (g,) = torch.autograd.grad(alpha_loss, [self.log_alpha])
        adam_step([self.log_alpha], [g], self.opt_alpha)
        self.alpha = self.log_alpha.detach().exp().item()
        self.scalars = dict(critic_loss=critic_loss.item(), actor_loss=actor_loss.item(),
                            alpha_loss=alpha_loss.item(), alpha=self.alpha,
                            logpi_mean=logp.mean().item())
        self.update_step += 1
        return self.scalars

    # ------------------------------------------------------------------- helpers
    def flat(self, which):
        groups = {
            "actor": self.actor,
            "critic": self._critic_flat(),
            "actor_target": self.actor_target or [],
            "critic_target": self._critic_target_flat(),
        }[which]
        if not groups:
            return np.zeros(0, np.float32)
        return torch.cat([p.detach().reshape(-1) for p in groups]).numpy().copy()


# --------------------------------------------------------------- parameter creation
def init_params(spec: AlgoSpec, seed: int):
    """Seeded synthetic initialisation with the reference's distributions' *scale*
    (nn.Linear default U(-1/sqrt(fan_in), 1/sqrt(fan_in))); used where goldens do not
    carry the reference's own initial weights (large TQC nets).  Stable across machines:
    numpy PCG64 uniform doubles -> fp32."""
    rng = np.random.default_rng(seed)

    def net(dims):
        out = []
        for shp in mlp_param_shapes(dims):
            fan_in = shp[1] if len(shp) == 2 else None
            if fan_in is None:
                fan_in = out[-1].shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            out.append(torch.from_numpy(rng.uniform(-bound, bound, size=shp).astype(np.float32)))
        return out

    actor = net(spec.actor_dims())
    critics = [net(spec.critic_dims()) for _ in range(spec.n_critics)]
    return actor, critics


# ------------------------------------------------------------------- replay buffer
def inds_to_episodic(inds: np.ndarray, ep_lens, episodes_counter: int):
    """EpisodicReplayBuffer._inds_to_episodic (episodic_buffer.py:114-121): transition index ->
    (episode, step) over the first ``episodes_counter`` episodes."""
    lens = np.asarray(ep_lens[:episodes_counter], dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int64)
    end = start + lens
    ep = np.argmin(inds.reshape(-1, 1) >= end.reshape(1, -1), axis=1)
    return ep, inds - start[ep]


def gather_batch(states, actions, rewards, dones, ep, step):
    """EpisodicReplayBuffer.sample after the index draw (episodic_buffer.py:127-133).
    states [E, L+1, S], actions [E, L, A], rewards/dones [E, L, 1] (numpy)."""
    return (states[ep, step], actions[ep, step], rewards[ep, step], dones[ep, step],
            states[ep, step + 1])


def synthetic_buffer(n_episodes, L, S, A, seed=0):
    """SURVEY.md section 8d synthetic replay content: zero-filled storage, episodes of exactly L
    steps, state ~ N(0,1), action ~ U(-1,1), reward ~ U(0,1), done = 0."""
    rng = np.random.default_rng(seed)
    states = np.zeros((n_episodes, L + 1, S), np.float32)
    states[:, :L] = rng.standard_normal((n_episodes, L, S), dtype=np.float32)
    actions = rng.uniform(-1, 1, (n_episodes, L, A)).astype(np.float32)
    rewards = rng.uniform(0, 1, (n_episodes, L, 1)).astype(np.float32)
    dones = np.zeros((n_episodes, L, 1), np.float32)
    return states, actions, rewards, dones


# ---------------------------------------------------------------------------- n-step returns (extension)
def nstep_batch(states, actions, rewards, dones, ep_lens, ep, step, n_step, gamma):
    """CPU statement of the engine's n-step gather (include/oprl_b200.h oprl_buffer_set_nstep).  The reference
    assembles 1-step transitions only (episodic_buffer.py:127-133; it stores `gamma`, :18, and never uses it), so
    there is nothing in the reference to pin this against: it is an EXTENSION defined here, and at n_step = 1 it
    must return exactly the reference's batch.  fp32 arithmetic in the kernel's order:
        R = r_t ; g = 1 ; for k = 1..m-1: g = g * gamma ; R = R + g * r_{t+k}        (m = window length)
        d' = 1 - (1 - d_{t+m-1}) * g ;  next_state = s_{t+m}
    with m = min(n_step, steps up to and including the first done, steps left in the episode)."""
    states, actions = np.asarray(states, np.float32), np.asarray(actions, np.float32)
    rewards, dones = np.asarray(rewards, np.float32), np.asarray(dones, np.float32)
    B = len(ep)
    S = states.shape[-1]
    out_s = states[ep, step].copy()
    out_a = actions[ep, step].copy()
    out_r = np.empty((B, 1), np.float32)
    out_d = np.empty((B, 1), np.float32)
    out_s2 = np.empty((B, S), np.float32)
    gam = np.float32(gamma)
    for i in range(B):
        e, t0 = int(ep[i]), int(step[i])
        R, d, g, m = np.float32(rewards[e, t0, 0]), np.float32(dones[e, t0, 0]), np.float32(1.0), 1
        while m < n_step and d == 0.0 and t0 + m < ep_lens[e]:
            g = np.float32(g * gam)
            R = np.float32(R + np.float32(g * rewards[e, t0 + m, 0]))
            d = np.float32(dones[e, t0 + m, 0])
            m += 1
        out_r[i, 0] = R
        out_d[i, 0] = np.float32(np.float32(1.0) - np.float32(np.float32(np.float32(1.0) - d) * g)) if n_step > 1 else d
        out_s2[i] = states[e, t0 + m]
    return out_s, out_a, out_r, out_d, out_s2
