"""GPU-resident episodic replay buffer behind the reference's ``EpisodicReplayBuffer`` surface
(episodic_buffer.py:13-140): same constructor fields, storage shapes
(states [E, L+1, S], actions [E, L, A], rewards / dones [E, L, 1]) and ring bookkeeping.

``sample()`` keeps the reference's host RNG stream (``np.random.randint`` on the global numpy
generator, episodic_buffer.py:124) and index -> (episode, step) mapping, then gathers on the
GPU: one CUDA kernel instead of five advanced-index ops.  When an algorithm attached its
engine (``algo.attach_buffer(buffer)``) the gather writes the engine's GEMM operand layout
directly and ``update()`` consumes it in place.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import numpy.typing as npt
import torch as t

from .. import _lib as L

Transition = tuple[npt.NDArray, npt.NDArray, float, bool, npt.NDArray]


def inds_to_episodic(inds: np.ndarray, ep_lens, episodes_counter: int):
    """Transition index -> (episode, step) over the first ``episodes_counter`` episodes; same
    result as the reference's dense compare + argmin (episodic_buffer.py:114-121) in
    O(B log E): the first episode whose end bound exceeds the index (0 if none does)."""
    lens = np.asarray(ep_lens[:episodes_counter], dtype=np.int64)
    end = np.cumsum(lens)
    start = end - lens
    ep = np.searchsorted(end, inds, side="right")
    ep[ep >= len(end)] = 0
    return ep, inds - start[ep]


@dataclass
class EpisodicReplayBuffer:
    buffer_size_transitions: int
    state_dim: int
    action_dim: int
    gamma: float = 0.99
    max_episode_lenth: int = 1000  # [sic] reference spelling, episodic_buffer.py:19
    episodes_counter: int = 1
    device: str = "cuda"
    # False (default): sample() returns fresh tensors like the reference's advanced indexing does.  True: it
    # returns views of the engine's one batch arena -- no copy, but the NEXT sample() overwrites them.
    # n-step returns assembled by the gather kernel (extension; needs an attached engine).  The reference stores
    # `gamma` here and never uses it (episodic_buffer.py:18): n_step = 1 keeps its 1-step transitions bit for bit.
    n_step = 1
    zero_copy_batches = False  # (plain class attribute: the dataclass fields stay exactly the reference's)

    _tensors: dict = field(init=False, default_factory=dict)
    _max_episodes: int = field(init=False, default=0)
    _ep_pointer: int = 0
    _number_transitions = 0
    _created: bool = False

    def create(self) -> "EpisodicReplayBuffer":
        dev = t.device(self.device)
        if dev.type != "cuda":
            raise L.EngineError("oprl_b200's replay buffer is GPU-resident: device must be CUDA")
        L.lib()  # fail loudly now if the CUDA library is missing
        E = self._max_episodes = self.buffer_size_transitions // self.max_episode_lenth
        Lmax = self.max_episode_lenth
        z = lambda *shape: t.zeros(shape, dtype=t.float32, device=dev)
        # zero-filled (the reference uses t.empty, whose never-written rows -- next_state of an
        # episode's last transition -- are indeterminate; zero is one admissible value)
        self._tensors = {"actions": z(E, Lmax, self.action_dim), "rewards": z(E, Lmax, 1),
                         "dones": z(E, Lmax, 1), "states": z(E, Lmax + 1, self.state_dim)}
        self.ep_lens = [0] * E
        self._stage_init()
        self._engine = None
        self._token = 0
        self._idx_dev = None
        self._created = True
        return self

    def check_created(self) -> None:
        if not self._created:
            raise RuntimeError("Replay buffer has to be created with `.create()`.")

    @property
    def states(self) -> t.Tensor:
        self.check_created()
        self.flush()  # readers see every transition added so far
        return self._tensors["states"]

    @property
    def actions(self) -> t.Tensor:
        self.check_created()
        self.flush()  # readers see every transition added so far
        return self._tensors["actions"]

    @property
    def rewards(self) -> t.Tensor:
        self.check_created()
        self.flush()  # readers see every transition added so far
        return self._tensors["rewards"]

    @property
    def dones(self) -> t.Tensor:
        self.check_created()
        self.flush()  # readers see every transition added so far
        return self._tensors["dones"]

    # ------------------------------------------------------------------ ingest edge
    # Transitions are staged in a pinned host ring ([state | action | reward | done | episode | step] per row) and
    # reach the replay storage in batches: ONE async H2D copy + ONE scatter kernel per flush -- before the next
    # sample(), when the ring is half full, and once per add_episode() -- instead of the two pageable copies and
    # two fill kernels per transition of a literal translation (episodic_buffer.py:89-96).
    _STAGE_ROWS = 4096

    def _stage_init(self) -> None:
        W = self.state_dim + self.action_dim + 4
        self._stage_host = t.zeros(self._STAGE_ROWS, W, dtype=t.float32).pin_memory()
        self._stage_np = self._stage_host.numpy()
        self._stage_int = self._stage_np.view(np.int32)
        self._stage_dev = t.zeros(self._STAGE_ROWS, W, dtype=t.float32, device=self.device)
        self._stage_lo = 0  # first row not yet flushed
        self._stage_hi = 0  # next free row
        self._stage_events = []  # (first row, last row + 1, event) of flushes whose H2D may still read the ring
        self._stage_eps = set()  # episode slots with staged, not yet flushed rows

    def _stage_wait_rows(self, lo: int, hi: int) -> None:
        keep = []
        for a, b, ev in self._stage_events:
            if a < hi and lo < b:
                ev.synchronize()
            elif not ev.query():
                keep.append((a, b, ev))
        self._stage_events = keep

    def flush(self) -> None:
        """Move every staged transition into the replay storage (async on the current stream)."""
        lo, hi = self._stage_lo, self._stage_hi
        if hi == lo or not self._created:
            return
        dev = self._stage_dev[lo:hi]
        dev.copy_(self._stage_host[lo:hi], non_blocking=True)
        T = self._tensors
        E, L1, S = T["states"].shape
        with t.cuda.device(T["states"].device):
            stream = t.cuda.current_stream(T["states"].device).cuda_stream or 1
            L.check(L.lib().oprl_scatter_transitions(
                T["states"].data_ptr(), T["actions"].data_ptr(), T["rewards"].data_ptr(), T["dones"].data_ptr(),
                E, L1 - 1, S, self.action_dim, dev.data_ptr(), hi - lo, C.c_void_p(stream)))
            ev = t.cuda.Event()
            ev.record()
        self._stage_events.append((lo, hi, ev))
        self._stage_eps.clear()
        if hi >= self._STAGE_ROWS:
            hi = 0
        self._stage_lo = self._stage_hi = hi

    def add_transition(self, state: npt.NDArray, action: npt.NDArray, reward: float, done: bool,
                       episode_done: bool | None = None) -> None:
        ep, i = self._ep_pointer, self.ep_lens[self._ep_pointer]
        if i >= self.max_episode_lenth:
            raise IndexError(f"episode {ep} exceeds max_episode_lenth={self.max_episode_lenth}")
        row = self._stage_hi
        if self._stage_events:
            self._stage_wait_rows(row, row + 1)
        S, A = self.state_dim, self.action_dim
        r = self._stage_np[row]
        r[:S] = state
        r[S:S + A] = action
        r[S + A] = reward
        r[S + A + 1] = float(done)
        ri = self._stage_int[row]
        ri[S + A + 2] = ep
        ri[S + A + 3] = i
        self._stage_eps.add(ep)
        self._stage_hi = row + 1
        if self._stage_hi - self._stage_lo >= self._STAGE_ROWS // 2 or self._stage_hi >= self._STAGE_ROWS:
            self.flush()
        self.ep_lens[ep] += 1
        self._number_transitions = min(self._number_transitions + 1, self.buffer_size_transitions)
        if episode_done:
            self._inc_episode()

    def _inc_episode(self) -> None:
        self._ep_pointer = (self._ep_pointer + 1) % self._max_episodes
        if self._ep_pointer in self._stage_eps:
            # the episode ring wrapped onto a slot that still has staged rows: one scatter launch writes its rows
            # in no particular order, so the older rows must land before newer ones for the same cells are staged
            self.flush()
        self.episodes_counter = min(self.episodes_counter + 1, self._max_episodes)
        self._number_transitions -= self.ep_lens[self._ep_pointer]
        self.ep_lens[self._ep_pointer] = 0

    def add_episode(self, episode: list[Transition]) -> None:
        # reference quirk kept: a terminal last transition advances the ring twice
        # (episodic_buffer.py:109-112)
        for s, a, r, d, _ in episode:
            self.add_transition(s, a, r, d, episode_done=d)
        self._inc_episode()
        self.flush()

    # --------------------------------------------------------------------- sample
    def attach_engine(self, engine) -> None:
        """Let ``sample()`` gather straight into ``engine``'s operand layout."""
        engine.bind_buffer(self.states, self.actions, self.rewards, self.dones)
        self._engine = engine

    def storage(self) -> dict:
        """The four storage tensors with every staged transition flushed (checkpoints, tests)."""
        self.flush()
        return self._tensors

    def draw_indices(self, batch_size: int) -> np.ndarray:
        """The reference's index draw: global numpy RNG, uniform over stored transitions."""
        inds = np.random.randint(low=0, high=self._number_transitions, size=batch_size)
        ep, step = inds_to_episodic(inds, self.ep_lens, self.episodes_counter)
        return np.stack([ep, step], axis=1).astype(np.int32)

    def sample(self, batch_size: int) -> tuple[t.Tensor, t.Tensor, t.Tensor, t.Tensor, t.Tensor]:
        self.flush()
        ep_step = self.draw_indices(batch_size)
        if self.n_step > 1:
            if self._engine is None:
                raise L.EngineError("n_step > 1 needs an attached engine (algo.attach_buffer(buffer))")
            if getattr(self, "_nstep_set", None) != (self.n_step, self.gamma):
                self._engine.set_nstep(self.n_step, self.gamma)
                self._nstep_set = (self.n_step, self.gamma)
            self._engine.set_prefix(self.ep_lens[:self.episodes_counter])  # the window stops at the episode's end
        if self._engine is not None:
            out = self._engine.sample(batch_size, ep_step)
            if not self.zero_copy_batches:
                out = self._engine.batch_copy(batch_size)
            self._token += 1
            token = (id(self), self._token)
            # The gather also wrote the engine's operand layout, so update(*batch) can skip its copy-in -- but
            # only for exactly these five tensors, unmodified, before another sample(): the token + version
            # stamp let update() detect anything else (a batch kept across a later sample(), replaced or edited
            # tensors) and fall back to loading what it was actually passed.
            for x in out:
                x._oprl_batch_token = token
                x._oprl_version = x._version
            self._engine.last_batch_token = token
            return out
        return self.gather(ep_step)

    def gather(self, ep_step: np.ndarray):
        """Engine-less gather of host-chosen (episode, step) pairs into fresh tensors."""
        B = ep_step.shape[0]
        S, A = self.state_dim, self.action_dim
        dev = self.states.device
        flat = t.empty(B * (2 * S + A + 2), dtype=t.float32, device=dev)
        o = 0
        outs = []
        for w in (S, A, 1, 1, S):
            outs.append(flat[o:o + B * w].view(B, w))
            o += B * w
        if self._idx_dev is None or self._idx_dev.numel() < 2 * B:
            self._idx_dev = t.empty(2 * B, dtype=t.int32, device=dev)
        ep_step = np.ascontiguousarray(ep_step, dtype=np.int32)
        E, L1, _ = self.states.shape
        with t.cuda.device(dev):
            stream = t.cuda.current_stream(dev).cuda_stream or 1
            L.check(L.lib().oprl_gather_rows(
                self.states.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(),
                self.dones.data_ptr(), E, L1 - 1, S, A, ep_step.ctypes.data,
                self._idx_dev.data_ptr(), B, *[x.data_ptr() for x in outs], C.c_void_p(stream)))
        return tuple(outs)

    @property
    def last_episode_length(self) -> int:
        return self.ep_lens[self._ep_pointer]

    def __len__(self) -> int:
        return self._number_transitions
