"""Shared helpers for the parity tests (fixtures -> oracle / engine state)."""
import os

import numpy as np
import torch

from oracle import oprl_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def spec_from_fixture(fx):
    algo = str(fx["algo"])
    return O.AlgoSpec(algo=algo, state_dim=int(fx["S"]), action_dim=int(fx["A"]),
                      tune_alpha=bool(int(fx["tune_alpha"])),
                      lr_alpha=3e-4 if algo == "tqc" else 1e-3)


def unflatten(flat, shapes):
    out, o = [], 0
    for shp in shapes:
        n = int(np.prod(shp))
        out.append(torch.from_numpy(np.asarray(flat[o:o + n], np.float32).reshape(shp).copy()))
        o += n
    assert o == len(flat)
    return out


def initial_params(fx):
    """(actor params, [critic net params]) the fixture was generated from."""
    spec = spec_from_fixture(fx)
    if int(fx["synthetic_init"]):
        return O.init_params(spec, int(fx["seed"]))
    actor = unflatten(fx["actor0"], O.mlp_param_shapes(spec.actor_dims()))
    cshapes = O.mlp_param_shapes(spec.critic_dims())
    per = sum(int(np.prod(s)) for s in cshapes)
    critics = [unflatten(fx["critic0"][i * per:(i + 1) * per], cshapes) for i in range(spec.n_critics)]
    return actor, critics


def oracle_from_fixture(fx):
    actor, critics = initial_params(fx)
    return O.OracleAlgo(spec_from_fixture(fx), actor, critics)


def fixture_batch(fx, k):
    return [torch.from_numpy(fx[f"{nm}{k}"]) for nm in ("s", "a", "r", "d", "s2")]


def fixture_noise(fx, k):
    out, i = [], 0
    while f"noise{k}_{i}" in fx:
        out.append(torch.from_numpy(fx[f"noise{k}_{i}"]))
        i += 1
    return out


def run_fixture_updates(orc, fx, k):
    return orc.update(*fixture_batch(fx, k), noise=fixture_noise(fx, k))
