"""Run the reference's own configs unchanged on the B200 engine.

``install()`` overlays the hot-path modules of an installed ``oprl`` package (schatty/oprl) with
their oprl_b200 counterparts, so that ``configs/ddpg.py`` & friends -- which import
``oprl.algos.ddpg.DDPG``, ``oprl.buffers.episodic_buffer.EpisodicReplayBuffer``, ... by name
(configs/ddpg.py:1-13) -- pick up the CUDA engine while everything outside the hot path
(environments, runners, logging, argument parsing) stays the reference's own code:

    python -m oprl_b200.compat configs/ddpg.py --env walker-walk --device cuda
"""
from __future__ import annotations

import importlib
import runpy
import sys

OVERLAY = {
    "oprl.algos.ddpg": "oprl_b200.algos.ddpg",
    "oprl.algos.td3": "oprl_b200.algos.td3",
    "oprl.algos.sac": "oprl_b200.algos.sac",
    "oprl.algos.tqc": "oprl_b200.algos.tqc",
    "oprl.algos.nn_models": "oprl_b200.algos.nn_models",
    "oprl.algos.nn_functions": "oprl_b200.algos.nn_functions",
    "oprl.buffers.episodic_buffer": "oprl_b200.buffers.episodic_buffer",
}


def install() -> list[str]:
    """Point the reference's hot-path module names at oprl_b200.  Returns the overlaid names.
    Must run before the config (or anything else) imports those modules."""
    done = []
    for ref_name, ours in OVERLAY.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition(".")
        if parent in sys.modules:  # keep `import oprl.algos; oprl.algos.ddpg` consistent
            setattr(sys.modules[parent], leaf, mod)
        done.append(ref_name)
    return done


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m oprl_b200.compat <config.py> [config args...]")
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
