"""Shared host logic of the four algorithms: arena adoption, batch hand-off, lazy metrics.

Reference surface: OffPolicyAlgorithm (base_algorithm.py:7-15) + the dataclass fields /
``create()`` / ``update()`` of ddpg.py, td3.py, sac.py, tqc.py.
"""
from __future__ import annotations

import os
from typing import Any

import torch as t
import torch.nn as nn

from ..engine import EngineSpec, UpdateEngine
from .nn_functions import disable_gradient
from .nn_models import adopt_parameters


class _DetachedHook:
    """What an ``_EngineDirtyHook`` unpickles to: a policy loaded from ``torch.save(algo.actor)`` has no engine."""

    def __call__(self, *_) -> None:
        return None


class _EngineDirtyHook:
    """``load_state_dict`` post-hook: the engine's tiled operand copies must be refreshed from theta.  A
    module-level class (not a closure) so that the reference's ``t.save(self.algo.actor, ...)``
    (trainers/base_trainer.py ``_save_policy``, distrib ``save_policy``) can pickle the module; the engine
    handle itself never travels."""

    def __init__(self, engine: UpdateEngine) -> None:
        self._engine = engine

    def __call__(self, *_) -> None:
        self._engine.mark_params_dirty()

    def __reduce__(self):
        return (_DetachedHook, ())


class EngineAdam:
    """Read-only view of the engine's fused Adam state (stands where the reference keeps a
    ``torch.optim.Adam``: ddpg.py:51,56)."""

    def __init__(self, engine: UpdateEngine, group: str, lr: float, step_field: str):
        self._engine, self._group, self.lr, self._step_field = engine, group, lr, step_field

    @property
    def exp_avg(self) -> t.Tensor:
        return self._engine.arena[self._group]["m"]

    @property
    def exp_avg_sq(self) -> t.Tensor:
        return self._engine.arena[self._group]["v"]

    @property
    def step_count(self) -> int:
        return int(getattr(self._engine.state(), self._step_field))

    def state_dict(self) -> dict:
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "step": self.step_count, "lr": self.lr}


class OffPolicyAlgorithm:
    _created: bool = False
    engine: UpdateEngine

    def check_created(self) -> None:
        if not self._created:
            raise RuntimeError(f"Algorithm {type(self).__name__} has not been created with `create()`.")

    def get_policy_state_dict(self) -> dict[str, Any]:
        return self.actor.state_dict()

    # ------------------------------------------------------------------ engine glue
    def _critic_nets(self) -> list[nn.Module]:
        raise NotImplementedError

    def _start_engine(self, spec: EngineSpec) -> None:
        # replicated learners must draw different minibatch rows / noise: offset the device RNG
        # streams by the process rank (torchrun's RANK)
        spec.seed = (spec.seed + 0x9E3779B1 * int(os.environ.get("RANK", "0"))) & 0xFFFFFFFFFFFFFFFF
        self.engine = UpdateEngine(spec, self.device)
        ar = self.engine.arena
        adopt_parameters(ar["actor"]["theta"], [self.actor])
        adopt_parameters(ar["critic"]["theta"], [self.critic])
        adopt_parameters(ar["critic"]["target"], [self.critic_target])
        disable_gradient(self.critic_target)
        if ar["actor"]["target"] is not None:
            adopt_parameters(ar["actor"]["target"], [self.actor_target])
            disable_gradient(self.actor_target)
        for mod in (self.actor, self.critic, self.critic_target, getattr(self, "actor_target", None)):
            if mod is not None:
                mod.register_load_state_dict_post_hook(_EngineDirtyHook(self.engine))
        self.engine.mark_params_dirty()

    def _hand_batch(self, state, action, reward, done, next_state) -> None:
        """update() accepts any five tensors (int64 ``done`` and aliased state / next_state
        included: tests/functional/test_rl_algos.py:25-31); tensors that are the engine's own
        sampled batch (``attach_buffer`` + ``sample``) are used in place."""
        five = (state, action, reward, done, next_state)
        token = getattr(state, "_oprl_batch_token", None)
        if token is not None and token == getattr(self.engine, "last_batch_token", None):
            # skip the copy-in only if ALL five tensors are the ones the latest sample() call handed out, with no
            # in-place edit since (tensor._version as stamped): the engine's operand layout then already holds them
            if all(getattr(x, "_oprl_batch_token", None) == token and x._version == getattr(x, "_oprl_version", -1)
                   for x in five):
                return
        self.engine.last_batch_token = None
        self.engine.load_batch(*five)

    # ------------------------------------------------------------------ host rollouts
    def enable_host_rollout(self, refresh_every: int = 1):
        """``actor.explore`` / ``actor.exploit`` run on a CPU mirror of the actor whose weights follow the device
        arena through asynchronous D2H copies into pinned memory every ``refresh_every`` updates -- the
        environment loop never synchronises with the GPU (reference acting path
        trainers/base_trainer.py:46-54)."""
        from .host_mirror import HostPolicyMirror

        if self.engine._params_dirty:
            self.engine.sync_params()
        self._mirror = HostPolicyMirror(self.actor, self.engine.arena["actor"]["theta"], refresh_every)
        self.actor.__dict__["_host_mirror"] = self._mirror
        return self._mirror

    def disable_host_rollout(self) -> None:
        self.actor.__dict__.pop("_host_mirror", None)
        self._mirror = None

    def attach_buffer(self, buffer) -> None:
        """Fuse ``buffer.sample()`` with this algorithm's engine: the gather kernel then writes
        the GEMM operand layout directly and ``update(*batch)`` skips the copy-in."""
        buffer.attach_engine(self.engine)

    # --------------------------------------------------------------- data parallel
    def enable_data_parallel(self, group=None, fused: bool = True) -> None:
        """Replicated learners (one process per GPU, identical parameters and replay content):
        every rank updates on its own B rows of a world_size * B minibatch and the flat gradient
        arenas are all-reduced (NCCL over NVLink) before each Adam step, so all replicas stay
        bit-identical without any parameter broadcast."""
        import torch.distributed as dist

        self._dp_group = group if group is not None else dist.group.WORLD
        self.engine.set_world_size(dist.get_world_size(self._dp_group))
        # start from rank 0's parameters (and targets); Adam state is zero everywhere at creation
        src = dist.get_global_rank(self._dp_group, 0)
        for grp in self.engine.arena.values():
            for key in ("theta", "target"):
                if grp[key] is not None:
                    dist.broadcast(grp[key], src=src, group=self._dp_group)
        self.engine.mark_params_dirty()
        if fused and os.environ.get("OPRL_B200_DP_NCCL", "0") != "1":
            # native path: peers' gradient arenas mapped over NVLink (CUDA IPC), the Adam kernel
            # sums them itself -- no NCCL launch between the segments
            rank = dist.get_rank(self._dp_group)
            world = dist.get_world_size(self._dp_group)
            mine = (self.engine.comm_init(rank, world), self.engine.device.index)
            gathered = [None] * world
            dist.all_gather_object(gathered, mine, group=self._dp_group)
            self.engine.comm_connect([h for h, _ in gathered], [d for _, d in gathered])
            dist.barrier(group=self._dp_group)

    def _run_update(self, actor_step: bool) -> None:
        self._run_update_device(actor_step)
        mirror = getattr(self, "_mirror", None)
        if mirror is not None:
            mirror.after_update()

    def _run_update_device(self, actor_step: bool) -> None:
        eng = self.engine
        group = getattr(self, "_dp_group", None)
        if group is None or eng.fused_comm:
            eng.update(actor_step=actor_step)
            return
        import torch.distributed as dist

        from .._lib import SEG_ACTOR_STEP, SEG_CRITIC_GRAD, SEG_CRITIC_STEP_ACTOR_GRAD

        eng.update(actor_step=actor_step, segment=SEG_CRITIC_GRAD)
        dist.all_reduce(eng.arena["critic"]["grad"], group=group)
        eng.update(actor_step=actor_step, segment=SEG_CRITIC_STEP_ACTOR_GRAD)
        if actor_step:
            dist.all_reduce(eng.arena["actor"]["grad"], group=group)
            eng.update(actor_step=actor_step, segment=SEG_ACTOR_STEP)

    def _wants_actor_step(self) -> bool:
        return True

    def learner_step(self, batch_size: int) -> None:
        """Device-resident learner iteration (extension; the reference's learner loop body
        ``batch = buffer.sample(B); algo.update(*batch)``, distrib/policy_update_worker.py:66-68,
        without the host round trip): uniform index draw on the GPU from the attached buffer,
        gather, update.  Needs ``attach_buffer`` + ``engine.set_prefix``."""
        eng = self.engine
        if getattr(self, "_dp_group", None) is None or eng.fused_comm:
            # one C call: in a steady loop the gather of the NEXT step already ran as a parallel branch
            # of this step's update graph (oprl_step), so nothing sits between two update graphs
            eng.step(batch_size, self._wants_actor_step())
        else:
            eng.sample(batch_size, None)
            self._run_update_device(self._wants_actor_step())
        mirror = getattr(self, "_mirror", None)
        if mirror is not None:
            mirror.after_update()
        self._after_update()

    def _after_update(self) -> None:
        self.update_step += 1

    def log_scalars_now(self) -> dict:
        """Device-side metrics of the last update (synchronises; only call on logging steps)."""
        return self.engine.scalars()
