"""Trainer / environment boundary (reference trainers/protocols.py:3-6, environment/protocols.py:6-34)."""
from typing import Any, Protocol

import numpy.typing as npt


class TrainerProtocol(Protocol):
    def train(self) -> None: ...

    def evaluate(self) -> dict[str, float]: ...


class EnvProtocol(Protocol):
    def step(self, action: npt.NDArray) -> tuple[npt.NDArray, float, bool, bool, dict[str, Any]]: ...

    def reset(self) -> tuple[npt.NDArray, dict[str, Any]]: ...

    def sample_action(self) -> npt.NDArray: ...
