"""Reference ``oprl.algos.nn_functions`` surface (nn_functions.py:5-16).  Inside the engine the
Polyak update is fused into the Adam kernel; these host versions serve user code that calls
them on its own modules."""
import torch as t
import torch.nn as nn


def soft_update(target: nn.Module, source: nn.Module, tau: float) -> None:
    with t.no_grad():
        for dst, src in zip(target.parameters(), source.parameters()):
            dst.data.mul_(1.0 - tau).add_(tau * src.data)


def disable_gradient(network: nn.Module) -> None:
    for p in network.parameters():
        p.requires_grad = False
