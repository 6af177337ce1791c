"""CPU-side mirror of the actor for rollouts ("next" row N1 of SURVEY.md section 8f).

The reference acts with the training network itself (``self.algo.actor.explore(state)``,
trainers/base_trainer.py:46-54; nn_models.py:144-150).  With the parameters living in the GPU arena that
is an H2D copy, three batch-1 GEMVs and a blocking D2H per environment step -- far longer than the
~50 us gradient update it sits next to.  The mirror keeps a second copy of the actor module on the
HOST whose parameters are views into pinned memory; every ``refresh_every`` updates the learner
enqueues one asynchronous D2H copy of the flat actor arena into the *back* pinned buffer (ordered
behind the update on the launch stream, no host wait), and the next ``explore`` / ``exploit`` call that
finds the copy complete swaps front and back.  Rollouts therefore run in host fp32 with weights at most
``refresh_every`` updates (plus one copy) old and never synchronise with the device.
"""
from __future__ import annotations

import copy

import numpy as np
import torch as t
import torch.nn as nn

LOG_STD_MIN_MAX = (-20.0, 2.0)  # nn_models.LOG_STD_MIN_MAX


class _NumpyMLP:
    """Batch-1 forward of an ``MLP`` on numpy views of the pinned parameter buffer: three small GEMVs cost ~10 us
    here against ~400 us through torch's CPU dispatcher and thread pool (measured on the GPU box: a torch CPU mirror
    was 3.6x SLOWER than acting on the device)."""

    def __init__(self, mlp: nn.Module):
        self.layers, self.acts = [], []
        mods = list(mlp.nn)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                act = mods[i + 1] if i + 1 < len(mods) else nn.Identity()
                if isinstance(act, nn.ReLU):
                    kind = "relu"
                elif isinstance(act, nn.Tanh):
                    kind = "tanh"
                elif isinstance(act, nn.Identity):
                    kind = "id"
                else:
                    raise TypeError(f"unsupported activation {type(act).__name__}")
                self.layers.append(m)
                self.acts.append(kind)
        self.bind()

    def bind(self) -> None:
        """(Re)take numpy views of the modules' current parameter storage (after a front / back buffer swap)."""
        self.wb = [(m.weight.data.numpy(), m.bias.data.numpy()) for m in self.layers]

    def __call__(self, x: np.ndarray) -> np.ndarray:
        h = np.asarray(x, dtype=np.float32).reshape(-1)
        for (w, b), kind in zip(self.wb, self.acts):
            h = w @ h + b
            if kind == "relu":
                np.maximum(h, 0.0, out=h)
            elif kind == "tanh":
                np.tanh(h, out=h)
        return h


class HostPolicyMirror:
    def __init__(self, actor: t.nn.Module, theta: t.Tensor, refresh_every: int = 1) -> None:
        self._theta = theta  # flat fp32 actor arena on the device (module.parameters() order)
        self.refresh_every = max(1, int(refresh_every))
        self._since = 0
        self._buf = [t.empty(theta.numel(), dtype=t.float32).pin_memory() for _ in range(2)]
        self._front = 0
        self._pending: t.cuda.Event | None = None
        # an independent CPU module of the same class; its parameters become views of the front buffer
        hooks = dict(actor._load_state_dict_post_hooks)
        actor._load_state_dict_post_hooks.clear()
        mirror_ref = actor.__dict__.pop("_host_mirror", None)
        try:
            self.module = copy.deepcopy(actor)
        finally:
            actor._load_state_dict_post_hooks.update(hooks)
            if mirror_ref is not None:
                actor.__dict__["_host_mirror"] = mirror_ref
        for attr in ("_device", "device"):
            if hasattr(self.module, attr):
                setattr(self.module, attr, "cpu")
        self._buf[0].copy_(theta)  # first fill: synchronous
        # numpy fast path for the two policy classes of the reference (nn_models.py:120-195); anything else acts
        # through the torch module
        self._np = None
        self._point_at(0)  # (first: the numpy views below are taken of the pinned front buffer)
        try:
            if hasattr(self.module, "mlp"):
                self._np = ("det", _NumpyMLP(self.module.mlp))
            elif hasattr(self.module, "net"):
                self._np = ("gauss", _NumpyMLP(self.module.net))
        except TypeError:  # an activation the fast path does not know
            self._np = None
        self.swaps = 0

    def _point_at(self, which: int) -> None:
        flat, off = self._buf[which], 0
        with t.no_grad():
            for p in self.module.parameters():
                n = p.numel()
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self._front = which
        if self._np is not None:
            self._np[1].bind()

    # ------------------------------------------------------------------ learner side
    def after_update(self) -> None:
        """Called once per gradient update: every ``refresh_every``-th call enqueues the async D2H."""
        self._since += 1
        if self._since < self.refresh_every or self._pending is not None:
            return
        self._since = 0
        back = self._front ^ 1
        self._buf[back].copy_(self._theta, non_blocking=True)
        ev = t.cuda.Event()
        ev.record()
        self._pending = ev

    def refresh_now(self) -> None:
        """Blocking refresh (evaluation, checkpoints)."""
        if self._pending is not None:
            self._pending.synchronize()
            self._pending = None
        back = self._front ^ 1
        self._buf[back].copy_(self._theta)
        self._point_at(back)
        self.swaps += 1

    # ------------------------------------------------------------------- rollout side
    def _maybe_swap(self) -> None:
        if self._pending is not None and self._pending.query():
            self._pending = None
            self._point_at(self._front ^ 1)
            self.swaps += 1

    def explore(self, state):
        self._maybe_swap()
        if self._np is None:
            return self.module.explore(state)
        kind, mlp = self._np
        m = self.module
        if kind == "det":
            # reference quirk kept: exploration acts on the pre-tanh output (nn_models.py:144-150); the noise comes
            # from torch's global CPU generator like the reference's t.randn
            noise = (t.randn(m._action_shape) * m._expl_noise).numpy()
            return np.clip(mlp(state) + noise, -m._max_action, m._max_action)
        out = mlp(state)
        A = m.action_dim
        mean, log_std = out[:A], out[A:]
        if not m.training:
            return np.tanh(mean)
        std = np.exp(np.clip(log_std, *LOG_STD_MIN_MAX))
        return np.tanh(mean + std * t.randn(A).numpy())

    def exploit(self, state):
        self._maybe_swap()
        if self._np is None:
            return self.module.exploit(state)
        kind, mlp = self._np
        out = mlp(state)
        if kind == "det":
            return np.tanh(out)
        return np.tanh(out[:self.module.action_dim])

    def __reduce__(self):  # never pickled with the policy (torch.save(algo.actor))
        return (_no_mirror, ())


def _no_mirror():
    return None
