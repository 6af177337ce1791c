"""Python handle of one CUDA update engine (include/oprl_b200.h).

PyTorch is used for device memory only: the flat parameter / gradient / Adam arenas and the
sampled-batch arena are torch CUDA tensors whose ``data_ptr()`` the engine borrows.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L


@dataclass
class EngineSpec:
    algo: str
    state_dim: int
    action_dim: int
    actor_hidden: int = 256
    actor_layers: int = 2
    critic_hidden: int = 256
    critic_layers: int = 2
    n_critics: int = 1
    n_quantiles: int = 1
    top_quantiles_to_drop: int = 0
    tune_alpha: bool = False
    gamma: float = 0.99
    tau: float = 5e-3
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    lr_alpha: float = 1e-3
    policy_noise: float = 0.2
    noise_clip: float = 0.5
    max_action: float = 1.0
    alpha_init: float = 0.2
    target_entropy: float = 0.0
    gemm_mode: int = L.GEMM_3XTF32
    world_size: int = 1
    seed: int = 0


class PendingScalars:
    """Ticket of one pipelined scalar read-back (UpdateEngine.scalars_async)."""

    def __init__(self, engine: "UpdateEngine", ticket: int):
        self._engine, self._ticket, self._out = engine, ticket, None

    def result(self) -> dict:
        if self._out is None:
            buf = (C.c_float * 32)()
            L.check(self._engine._lib.oprl_scalars_wait(self._engine._h, self._ticket, buf, 32))
            self._out = {k: float(buf[i]) for i, k in enumerate(L.SCALARS)}
        return self._out


class UpdateEngine:
    """Owns the engine handle and the torch tensors it borrows."""

    def __init__(self, spec: EngineSpec, device: str | torch.device = "cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.EngineError("oprl_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if not torch.cuda.is_available():
            raise L.EngineError("no CUDA device visible: oprl_b200 has no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        if "OPRL_B200_GEMM_MODE" in os.environ:  # 0 = 3xTF32 (default), 1 = single TF32, 2 = FFMA cross-check
            spec = dataclasses.replace(spec, gemm_mode=int(os.environ["OPRL_B200_GEMM_MODE"]))
        self.spec = spec
        self._lib = L.lib()
        cfg = L.Cfg(
            algo=L.ALGO[spec.algo], state_dim=spec.state_dim, action_dim=spec.action_dim,
            actor_hidden=spec.actor_hidden, actor_layers=spec.actor_layers,
            critic_hidden=spec.critic_hidden, critic_layers=spec.critic_layers,
            n_critics=spec.n_critics, n_quantiles=spec.n_quantiles,
            top_quantiles_to_drop=spec.top_quantiles_to_drop, tune_alpha=int(spec.tune_alpha),
            gemm_mode=spec.gemm_mode, device=self.device.index, world_size=spec.world_size,
            gamma=spec.gamma, tau=spec.tau, lr_actor=spec.lr_actor, lr_critic=spec.lr_critic,
            lr_alpha=spec.lr_alpha, policy_noise=spec.policy_noise, noise_clip=spec.noise_clip,
            max_action=spec.max_action, alpha_init=spec.alpha_init,
            target_entropy=spec.target_entropy, seed=spec.seed)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self._lib.oprl_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._stream_id = None
        self.arena = {}
        for net, name in ((L.NET_ACTOR, "actor"), (L.NET_CRITIC, "critic")):
            n = int(self._lib.oprl_engine_arena_floats(h, net))
            z = lambda extra=0: torch.zeros(n + extra, dtype=torch.float32, device=self.device)
            has_target = name == "critic" or spec.algo in ("ddpg", "td3")
            a = dict(theta=z(), grad=z(L.GRAD_TAIL), m=z(), v=z(), target=z() if has_target else None)
            self.arena[name] = a
            L.check(self._lib.oprl_engine_bind_arena(
                h, net, a["theta"].data_ptr(), a["grad"].data_ptr(), a["m"].data_ptr(),
                a["v"].data_ptr(), a["target"].data_ptr() if has_target else None))
        self._batch = None
        self._batch_cap = 0
        self._scalars = (C.c_float * 32)()
        self._params_dirty = True
        self._keep = []  # tensors the engine holds pointers into
        self.last_batch_token = None
        self.fused_comm = False

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.oprl_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------------- streams
    def _use_current_stream(self):
        # 0 is torch's legacy default stream: name it explicitly (cudaStreamLegacy == 0x1), NULL
        # would mean "engine-owned stream" to the C ABI.
        s = torch._C._cuda_getCurrentRawStream(self.device.index) or 1
        if s != self._stream_id:
            L.check(self._lib.oprl_engine_set_stream(self._h, C.c_void_p(s)))
            self._stream_id = s

    # ----------------------------------------------------------------- parameters
    def mark_params_dirty(self):
        self._params_dirty = True

    def sync_params(self):
        """Re-tile the tf32 operand copies from theta / theta_target (after load_state_dict)."""
        self._use_current_stream()
        L.check(self._lib.oprl_engine_sync_params(self._h))
        self._params_dirty = False

    # --------------------------------------------------------------------- replay
    def bind_buffer(self, states, actions, rewards, dones):
        E, L1, S = states.shape
        for t in (states, actions, rewards, dones):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        L.check(self._lib.oprl_buffer_bind(self._h, states.data_ptr(), actions.data_ptr(),
                                           rewards.data_ptr(), dones.data_ptr(), E, L1 - 1))
        self._keep.append((states, actions, rewards, dones))

    def set_prefix(self, ep_lens):
        pre = np.zeros(len(ep_lens) + 1, np.int32)
        np.cumsum(np.asarray(ep_lens, np.int64), out=pre[1:])
        L.check(self._lib.oprl_buffer_set_prefix(self._h, pre.ctypes.data, len(ep_lens)))

    def set_nstep(self, n_step: int, gamma: float):
        """n-step return assembly in the gather (extension; 1 = the reference's 1-step transitions)."""
        L.check(self._lib.oprl_buffer_set_nstep(self._h, int(n_step), float(gamma)))

    def _ensure_batch(self, B):
        if self._batch is None or self._batch_cap < B:
            cap = max(B, 128)
            sp = self.spec
            widths = (sp.state_dim, sp.action_dim, 1, 1, sp.state_dim)
            # one flat allocation, five row-major [cap, w] views: a caller-visible copy of a sampled batch is one
            # D2D copy of `flat` (EpisodicReplayBuffer.sample)
            self._batch_flat = torch.zeros(cap * sum(widths), dtype=torch.float32, device=self.device)
            views, o = [], 0
            for w in widths:
                views.append(self._batch_flat[o:o + cap * w].view(cap, w))
                o += cap * w
            self._batch = tuple(views)
            self._batch_cap = cap
            L.check(self._lib.oprl_batch_bind(self._h, *[t.data_ptr() for t in self._batch], cap))
        return self._batch

    def batch_views(self, B):
        return tuple(t[:B] for t in self._ensure_batch(B))

    def batch_copy(self, B):
        """Fresh tensors holding the batch the last sample() gathered (one D2D copy of the flat arena)."""
        self._ensure_batch(B)
        flat = self._batch_flat.clone()
        sp = self.spec
        cap, out, o = self._batch_cap, [], 0
        for w in (sp.state_dim, sp.action_dim, 1, 1, sp.state_dim):
            out.append(flat[o:o + cap * w].view(cap, w)[:B])
            o += cap * w
        return tuple(out)

    def sample(self, B, ep_step: np.ndarray | None = None):
        """Gather B transitions straight from the bound replay storage into the engine's operand
        layout.  ep_step: int32 [B, 2] host array (reference index parity) or None = device RNG."""
        self._ensure_batch(B)
        self._use_current_stream()
        if ep_step is not None:
            ep_step = np.ascontiguousarray(ep_step, dtype=np.int32)
            assert ep_step.shape == (B, 2)
            ptr = ep_step.ctypes.data
        else:
            ptr = None
        L.check(self._lib.oprl_sample(self._h, ptr, B))
        return self.batch_views(B)

    def load_batch(self, s, a, r, d, s2):
        B = s.shape[0]
        self._ensure_batch(B)
        self._use_current_stream()
        if not s.is_cuda:
            # host batch: one packed pinned staging copy inside the engine
            hs = []
            for t, w in ((s, self.spec.state_dim), (a, self.spec.action_dim), (r, 1), (d, 1),
                         (s2, self.spec.state_dim)):
                if t.is_cuda:
                    t = t.cpu()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.to(torch.float32).contiguous()
                assert t.numel() == B * w, "batch tensor has the wrong shape"
                hs.append(t)
            L.check(self._lib.oprl_load_batch_host(self._h, *[t.data_ptr() for t in hs], B))
            return
        ts = []
        for t, w in ((s, self.spec.state_dim), (a, self.spec.action_dim), (r, 1), (d, 1),
                     (s2, self.spec.state_dim)):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
            assert t.numel() == B * w, "batch tensor has the wrong shape"
            ts.append(t)
        L.check(self._lib.oprl_load_batch(self._h, *[t.data_ptr() for t in ts], B))
        self._last_inputs = ts  # keep alive until the next call

    def set_noise(self, which, noise):
        self._use_current_stream()
        noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        L.check(self._lib.oprl_set_noise(self._h, which, noise.data_ptr(), noise.numel()))
        self._last_noise = noise

    # --------------------------------------------------------------------- update
    def update(self, actor_step=True, segment=L.SEG_ALL):
        if self._params_dirty:
            self.sync_params()
        self._use_current_stream()
        L.check(self._lib.oprl_update(self._h, L.UPDATE_ACTOR if actor_step else 0, segment))

    def step(self, B, actor_step=True):
        """Device-resident learner step: on-device uniform sampling + update."""
        if self._params_dirty:
            self.sync_params()
        self._ensure_batch(B)
        self._use_current_stream()
        L.check(self._lib.oprl_step(self._h, B, L.UPDATE_ACTOR if actor_step else 0))

    def set_world_size(self, world_size: int):
        L.check(self._lib.oprl_engine_set_world_size(self._h, int(world_size)))

    def comm_init(self, rank: int, world: int) -> bytes:
        """This rank's three CUDA-IPC handles (actor grads, critic grads, flags)."""
        buf = C.create_string_buffer(3 * 64)
        L.check(self._lib.oprl_comm_init(self._h, rank, world, buf))
        return buf.raw

    def comm_connect(self, handles: list, devices: list):
        blob = b"".join(handles)
        dev = (C.c_int * len(devices))(*devices)
        L.check(self._lib.oprl_comm_connect(self._h, blob, dev))
        self.fused_comm = True

    def launches(self, B, actor_step=True):
        return L.check(self._lib.oprl_update_launches(self._h, B, L.UPDATE_ACTOR if actor_step else 0))

    # ------------------------------------------------------------------ profiling
    def time_gemm_only(self, B, iters=200, actor_step=True):
        """(ms per update spent in the update's GEMM launches, GEMM launches per update)."""
        self._use_current_stream()
        ms, n = C.c_float(), C.c_int()
        L.check(self._lib.oprl_profile(self._h, B, L.UPDATE_ACTOR if actor_step else 0, 0, iters,
                                       C.byref(ms), C.byref(n)))
        return ms.value / iters, n.value

    def time_simt_only(self, B, iters=200, actor_step=True):
        """(ms per update spent in the update's non-GEMM launches, launches per update)."""
        self._use_current_stream()
        ms, n = C.c_float(), C.c_int()
        L.check(self._lib.oprl_profile(self._h, B, L.UPDATE_ACTOR if actor_step else 0, 2, iters,
                                       C.byref(ms), C.byref(n)))
        return ms.value / iters

    def time_chain_only(self, B, iters=200, actor_step=True):
        """(ms per update spent in the update's batch-slice chain launches, chain launches per update)."""
        self._use_current_stream()
        ms, n = C.c_float(), C.c_int()
        L.check(self._lib.oprl_profile(self._h, B, L.UPDATE_ACTOR if actor_step else 0, 3, iters,
                                       C.byref(ms), C.byref(n)))
        return ms.value / iters, n.value

    def chain_prof(self, B, which, actor_step=True):
        """clock64 stamps of CTA 0 of the critic (0) / actor (1) chain launch (OPRL_B200_CHAIN_PROF=1)."""
        out = (C.c_longlong * 256)()
        L.check(self._lib.oprl_chain_prof(self._h, B, L.UPDATE_ACTOR if actor_step else 0, which, out))
        return list(out)

    def time_gather_only(self, B, iters=200):
        """Microseconds per gather launch (device-side index draw)."""
        self._use_current_stream()
        ms, n = C.c_float(), C.c_int()
        L.check(self._lib.oprl_profile(self._h, B, L.UPDATE_ACTOR, 1, iters, C.byref(ms), C.byref(n)))
        return ms.value / iters * 1e3

    # -------------------------------------------------------------------- scalars
    def scalars(self) -> dict:
        L.check(self._lib.oprl_get_scalars(self._h, self._scalars, 32))
        return {k: float(self._scalars[i]) for i, k in enumerate(L.SCALARS)}

    def scalars_async(self) -> "PendingScalars":
        """Enqueue a D2H read of the scalars of everything launched so far; `.result()` waits for
        that copy only, so the next update can already be running (valid for 8 further calls)."""
        self._use_current_stream()
        return PendingScalars(self, L.check(self._lib.oprl_scalars_enqueue(self._h)))

    def state(self) -> L.State:
        st = L.State()
        L.check(self._lib.oprl_get_state(self._h, C.byref(st)))
        return st

    def set_state(self, **kw):
        st = self.state()
        for k, v in kw.items():
            setattr(st, k, v)
        L.check(self._lib.oprl_set_state(self._h, C.byref(st)))

    def sync(self):
        L.check(self._lib.oprl_sync(self._h))
