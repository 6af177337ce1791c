"""CPU: the C-ABI shared library loads and exports every symbol include/oprl_b200.h declares;
entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oprl_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oprl_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from oprl_b200 import _lib

    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/oprl_b200.h but not exported"


def test_python_binding_covers_the_header(lib):
    from oprl_b200 import _lib

    assert set(declared_functions()) == set(_lib._SIGNATURES)


def test_abi_version_and_struct_layout(lib):
    from oprl_b200 import _lib

    assert lib.oprl_abi_version() == 1
    assert C.sizeof(_lib.Cfg) == 14 * 4 + 10 * 8 + 8
    assert C.sizeof(_lib.State) == 8 + 4 * 4 + 3 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu(lib):
    from oprl_b200 import _lib

    cfg = _lib.Cfg(algo=0, state_dim=24, action_dim=6, actor_hidden=256, actor_layers=2,
                   critic_hidden=256, critic_layers=2, n_critics=1, n_quantiles=1, world_size=1)
    h = C.c_void_p()
    rc = lib.oprl_engine_create(C.byref(cfg), C.byref(h))
    assert rc < 0 and not h
    assert b"no CPU fallback" in lib.oprl_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_algorithms_refuse_cpu():
    from oprl_b200._lib import EngineError
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    with pytest.raises(EngineError):
        DDPG(logger=None, state_dim=24, action_dim=6, device="cpu").create()
    with pytest.raises(EngineError):
        DDPG(logger=None, state_dim=24, action_dim=6, device="cuda").create()
    with pytest.raises(EngineError):
        EpisodicReplayBuffer(1000, 3, 2, device="cpu").create()


def test_bad_arguments_are_reported(lib):
    assert lib.oprl_engine_create(None, None) < 0
    assert lib.oprl_last_error()
    assert lib.oprl_update(None, 0, -1) < 0
    assert lib.oprl_engine_arena_floats(None, 0) == -1


def test_product_code_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "oprl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oprl_oracle" in txt:
                    bad.append(f)
    assert not bad, bad
