"""Actor <-> learner transport without a message broker ("next" row N2).

The reference moves pickled episodes and policy weights through RabbitMQ (distrib/queue.py:4-19,
``Queue(name, host).push(bytes) / .pop() -> bytes | None``) and polls with 1-2 s sleeps -- far
too coarse for a learner whose update takes ~0.1 ms.  Here the same ``Queue`` surface is served
by an in-node ``multiprocessing`` manager: named FIFO queues, plus a blocking ``pop_wait`` so
neither side sleeps on a timer.
"""
from __future__ import annotations

import os
import queue as _queue
from multiprocessing.managers import BaseManager

_AUTH = b"oprl_b200"
_registry: dict[str, _queue.Queue] = {}


def _get(name: str) -> _queue.Queue:
    return _registry.setdefault(name, _queue.Queue())


class _Manager(BaseManager):
    pass


_Manager.register("get_queue", callable=_get)


def _port() -> int:
    return int(os.environ.get("OPRL_B200_QUEUE_PORT", "56721"))


class QueueServer:
    """Owns the named queues; started by the process that spawns the workers."""

    def __init__(self, host: str = "127.0.0.1", port: int | None = None) -> None:
        self._mgr = _Manager(address=(host, port or _port()), authkey=_AUTH)

    def __enter__(self) -> "QueueServer":
        self._mgr.start()
        return self

    def __exit__(self, *exc) -> None:
        self._mgr.shutdown()


class Queue:
    def __init__(self, name: str, host: str = "localhost") -> None:
        self._name = name
        mgr = _Manager(address=("127.0.0.1" if host == "localhost" else host, _port()), authkey=_AUTH)
        mgr.connect()
        self._q = mgr.get_queue(name)

    def push(self, data) -> None:
        self._q.put(data)

    def pop(self) -> bytes | None:
        """Non-blocking, like the reference's ``basic_get``: ``None`` when the queue is empty."""
        try:
            return self._q.get_nowait()
        except _queue.Empty:
            return None

    def pop_wait(self, timeout: float) -> bytes | None:
        """Block up to ``timeout`` seconds for the next message."""
        try:
            return self._q.get(True, timeout)
        except _queue.Empty:
            return None
