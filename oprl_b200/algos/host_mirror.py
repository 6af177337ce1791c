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

import torch as t


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
        self._point_at(0)
        self.swaps = 0

    def _point_at(self, which: int) -> None:
        flat, off = self._buf[which], 0
        with t.no_grad():
            for p in self.module.parameters():
                n = p.numel()
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self._front = which

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
        return self.module.explore(state)

    def exploit(self, state):
        self._maybe_swap()
        return self.module.exploit(state)

    def __reduce__(self):  # never pickled with the policy (torch.save(algo.actor))
        return (_no_mirror, ())


def _no_mirror():
    return None
