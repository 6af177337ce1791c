"""Full-state checkpoint / resume of a learner ("next" row N3): parameter, target and Adam arenas,
optimizer step counts, RNG tick, temperature state, and optionally the replay buffer.  The
reference saves only ``t.save(actor)`` (base_trainer.py:113-120) and has no resume path."""
from __future__ import annotations

from typing import Any

import torch


def algo_state(algo) -> dict[str, Any]:
    eng = algo.engine
    st = eng.state()
    return {
        "format": 1,
        "algo": type(algo).__name__,
        "update_step": int(getattr(algo, "update_step", 0)),
        "arenas": {grp: {k: v.detach().cpu().clone() for k, v in a.items() if v is not None and k != "grad"}
                   for grp, a in eng.arena.items()},
        "engine": {f: getattr(st, f) for f in ("tick", "step_actor", "step_critic", "step_alpha",
                                               "log_alpha", "m_alpha", "v_alpha")},
    }


def load_algo_state(algo, state: dict[str, Any]) -> None:
    if state.get("format") != 1 or state.get("algo") != type(algo).__name__:
        raise ValueError("checkpoint does not match this algorithm")
    eng = algo.engine
    for grp, tensors in state["arenas"].items():
        for k, v in tensors.items():
            dst = eng.arena[grp][k]
            if dst.numel() != v.numel():
                raise ValueError(f"arena {grp}/{k}: size mismatch")
            dst.copy_(v.to(dst.device))
    eng.set_state(**state["engine"])
    eng.mark_params_dirty()
    if hasattr(algo, "update_step"):
        algo.update_step = state["update_step"]


def buffer_state(buf) -> dict[str, Any]:
    return {
        "tensors": {k: v.detach().cpu().clone() for k, v in buf.storage().items()},
        "ep_lens": list(buf.ep_lens),
        "ep_pointer": buf._ep_pointer,
        "episodes_counter": buf.episodes_counter,
        "number_transitions": buf._number_transitions,
    }


def load_buffer_state(buf, state: dict[str, Any]) -> None:
    buf.flush()
    for k, v in state["tensors"].items():
        buf._tensors[k].copy_(v.to(buf._tensors[k].device))
    buf.ep_lens = list(state["ep_lens"])
    buf._ep_pointer = state["ep_pointer"]
    buf.episodes_counter = state["episodes_counter"]
    buf._number_transitions = state["number_transitions"]


def save_checkpoint(path, algo, replay_buffer=None) -> None:
    blob = {"algo": algo_state(algo)}
    if replay_buffer is not None:
        blob["buffer"] = buffer_state(replay_buffer)
    torch.save(blob, path)


def load_checkpoint(path, algo, replay_buffer=None) -> None:
    blob = torch.load(path, map_location="cpu", weights_only=True)  # tensors + plain containers only
    load_algo_state(algo, blob["algo"])
    if replay_buffer is not None and "buffer" in blob:
        load_buffer_state(replay_buffer, blob["buffer"])
