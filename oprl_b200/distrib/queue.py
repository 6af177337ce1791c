"""Actor <-> learner transport without a message broker ("next" row N2).

The reference moves pickled episodes and policy weights through RabbitMQ (distrib/queue.py:4-19,
``Queue(name, host).push(bytes) / .pop() -> bytes | None``) and polls with 1-2 s sleeps -- far too coarse
for a learner whose update takes ~0.05 ms.  Here the same ``Queue`` surface is a single-producer /
single-consumer byte ring in POSIX shared memory: no server process, no socket, no authentication key
to leak -- the segment is private to the user (mode 0600 under /dev/shm) and its name carries a random
per-run token (``OPRL_B200_SESSION``, set by ``QueueServer`` and inherited by the workers it spawns).
Episodes and weights travel as raw float32 (``pack_episode`` / ``pack_weights``): nothing on this path is
unpickled.

Layout of a ring: 64-byte header (``head``, ``tail`` as int64 byte counters that only grow, capacity) +
``capacity`` data bytes; a message is a 8-byte length prefix + payload, wrapped around the end.
"""
from __future__ import annotations

import os
import secrets
import struct
import time
from multiprocessing import shared_memory

import numpy as np

_HDR = 64
_DEFAULT_CAP = 64 << 20  # bytes per ring: hundreds of 1000-step episodes, or a few policy snapshots


def _session() -> str:
    tok = os.environ.get("OPRL_B200_SESSION")
    if not tok:
        raise RuntimeError("no queue session: start the workers under `with QueueServer():` (it sets OPRL_B200_SESSION)")
    return tok


def _shm_name(name: str) -> str:
    safe = "".join(ch if ch.isalnum() else "_" for ch in name)
    return f"oprl_b200_{_session()}_{safe}"


class QueueServer:
    """Owns the rings of one training run: creates them on first use by name and unlinks them on exit.
    ``names`` pre-creates rings (the spawning process knows every queue of the topology)."""

    def __init__(self, names: list[str] | None = None, capacity: int = _DEFAULT_CAP) -> None:
        self._names = list(names or [])
        self._cap = capacity
        self._owned: list[shared_memory.SharedMemory] = []

    def __enter__(self) -> "QueueServer":
        os.environ["OPRL_B200_SESSION"] = secrets.token_hex(8)
        for n in self._names:
            self.create(n)
        return self

    def create(self, name: str) -> None:
        shm = shared_memory.SharedMemory(name=_shm_name(name), create=True, size=_HDR + self._cap)
        hdr = np.ndarray(8, dtype=np.int64, buffer=shm.buf)
        hdr[:] = 0
        hdr[2] = self._cap
        self._owned.append(shm)

    def __exit__(self, *exc) -> None:
        for shm in self._owned:
            try:
                shm.close()
                shm.unlink()
            except FileNotFoundError:
                pass
        self._owned.clear()
        os.environ.pop("OPRL_B200_SESSION", None)


class Queue:
    """``push(bytes)`` / ``pop() -> bytes | None`` like the reference's queue, plus a blocking ``pop_wait``.
    One producer process and one consumer process per queue."""

    def __init__(self, name: str, host: str = "localhost") -> None:
        if host not in ("localhost", "127.0.0.1"):
            raise ValueError("oprl_b200 queues are in-node shared memory: host must be localhost")
        self._name = name
        self._shm = shared_memory.SharedMemory(name=_shm_name(name))
        # (python 3.12 registers the segment with the resource tracker on attach as well; the workers are spawned by
        # the process that owns the QueueServer and share its tracker, whose name cache is a set -- the owner's
        # unlink() is the one removal)
        self._hdr = np.ndarray(8, dtype=np.int64, buffer=self._shm.buf)
        self._cap = int(self._hdr[2])
        self._data = np.ndarray(self._cap, dtype=np.uint8, buffer=self._shm.buf, offset=_HDR)

    def close(self) -> None:
        self._hdr = self._data = None
        self._shm.close()

    # ring primitives ---------------------------------------------------------------------
    def _write(self, pos: int, raw: np.ndarray) -> None:
        off, n = pos % self._cap, raw.size
        first = min(n, self._cap - off)
        self._data[off:off + first] = raw[:first]
        if first < n:
            self._data[:n - first] = raw[first:]

    def _read(self, pos: int, n: int) -> np.ndarray:
        off = pos % self._cap
        first = min(n, self._cap - off)
        out = np.empty(n, np.uint8)
        out[:first] = self._data[off:off + first]
        if first < n:
            out[first:] = self._data[:n - first]
        return out

    def push(self, data) -> None:
        raw = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)
        need = 8 + raw.size
        if need > self._cap:
            raise ValueError(f"message of {raw.size} bytes exceeds the ring capacity {self._cap}")
        head = int(self._hdr[0])
        delay = 50e-6
        while head + need - int(self._hdr[1]) > self._cap:  # wait for the consumer to free space
            time.sleep(delay)
            delay = min(delay * 2, 2e-3)
        self._write(head, np.frombuffer(struct.pack("<q", raw.size), dtype=np.uint8))
        self._write(head + 8, raw)
        self._hdr[0] = head + need  # publish (x86-TSO: the payload stores above are visible before this one)

    def pop(self) -> bytes | None:
        """Non-blocking, like the reference's ``basic_get``: ``None`` when the queue is empty."""
        tail = int(self._hdr[1])
        if int(self._hdr[0]) == tail:
            return None
        (n,) = struct.unpack("<q", self._read(tail, 8).tobytes())
        out = self._read(tail + 8, n).tobytes()
        self._hdr[1] = tail + 8 + n
        return out

    def pop_wait(self, timeout: float) -> bytes | None:
        """Block up to ``timeout`` seconds for the next message (short exponential back-off polls)."""
        t_end = time.monotonic() + timeout
        delay = 20e-6
        while True:
            out = self.pop()
            if out is not None or time.monotonic() >= t_end:
                return out
            time.sleep(delay)
            delay = min(delay * 2, 1e-3)


# ------------------------------------------------------------------------------ message formats
_EP_MAGIC, _W_MAGIC = 0x4F50524C45503031, 0x4F50524C57543031  # "OPRLEP01" / "OPRLWT01"


def pack_episode(episode, state_dim: int, action_dim: int) -> bytes:
    """[state, action, reward, terminated, next_state] per step (reference env_worker.py:36) -> raw float32
    rows [state | action | reward | done | next_state]."""
    T, W = len(episode), 2 * state_dim + action_dim + 2
    arr = np.empty((T, W), np.float32)
    for i, (s, a, r, d, s2) in enumerate(episode):
        arr[i, :state_dim] = s
        arr[i, state_dim:state_dim + action_dim] = a
        arr[i, state_dim + action_dim] = r
        arr[i, state_dim + action_dim + 1] = float(d)
        arr[i, state_dim + action_dim + 2:] = s2
    return struct.pack("<qqqq", _EP_MAGIC, T, state_dim, action_dim) + arr.tobytes()


def unpack_episode(data: bytes) -> tuple[np.ndarray, int, int]:
    """-> (rows [T, 2S + A + 2], S, A)."""
    magic, T, S, A = struct.unpack_from("<qqqq", data, 0)
    if magic != _EP_MAGIC:
        raise ValueError("not an episode message")
    arr = np.frombuffer(data, dtype=np.float32, offset=32).reshape(T, 2 * S + A + 2)
    return arr, S, A


def episode_rows_to_list(arr: np.ndarray, S: int, A: int):
    """Back to the reference's list-of-transitions form (what ``add_episode`` takes)."""
    return [(r[:S], r[S:S + A], float(r[S + A]), bool(r[S + A + 1]), r[S + A + 2:]) for r in arr]


def pack_weights(state_dict) -> bytes:
    """Policy ``state_dict`` -> names + raw float32 (reference ships ``pickle.dumps(state_dict)``,
    policy_update_worker.py:74-76)."""
    names, shapes, blobs = [], [], []
    for k, v in state_dict.items():
        a = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
        names.append(k)
        shapes.append(tuple(a.shape))  # (ascontiguousarray would turn a 0-d tensor into shape (1,))
        blobs.append(a.tobytes())
    meta = repr((names, shapes)).encode()
    return struct.pack("<qq", _W_MAGIC, len(meta)) + meta + b"".join(blobs)


def unpack_weights(data: bytes):
    import ast

    import torch

    magic, n_meta = struct.unpack_from("<qq", data, 0)
    if magic != _W_MAGIC:
        raise ValueError("not a weights message")
    names, shapes = ast.literal_eval(data[16:16 + n_meta].decode())
    out, off = {}, 16 + n_meta
    for k, shp in zip(names, shapes):
        n = int(np.prod(shp)) if len(shp) else 1
        out[k] = torch.from_numpy(np.frombuffer(data, dtype=np.float32, count=n, offset=off).reshape(shp).copy())
        off += 4 * n
    return out
