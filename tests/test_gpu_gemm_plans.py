"""GPU: the launch-plan choices of the grouped GEMM path keep the parity bar whichever way they fall.

By default `prepare_stage_tables` (csrc/engine.cu) gives launches that exceed one wave of SMs 128 x 64 tiles, may issue
a stage as two launches when its cost model says so, re-tiles large weight matrices in 32 x 32 patches inside the
Adam kernel (TQC), and leaves the sum over M tiles of the fused layer-0 gradient to the Adam kernel.  The SAC (batch 1024) and TQC fixtures run through those paths in tests/test_gpu_parity.py; here
the same fixtures run with each choice switched off in a child interpreter (the switches are read once per process),
and the plans are checked to differ the way the switches say.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PLAN_PROBE = """
import sys
sys.path.insert(0, %r)
from bench import make_algo, WORKLOADS
wl = WORKLOADS["tqc"]
a = make_algo("tqc", wl["S"], wl["A"], "cuda:0")
print("LAUNCHES", a.engine.launches(wl["B"], True))
""" % ROOT


def _plan(env):
    child_env = dict(os.environ, OPRL_B200_DUMP_STAGES="1", **env)
    out = subprocess.run([sys.executable, "-c", PLAN_PROBE], env=child_env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stderr


def test_wide_tiles_are_used_where_a_launch_exceeds_one_wave():
    plan = _plan({})
    assert "(128 x 64)" in plan, plan[-1500:]
    narrow = _plan({"OPRL_B200_GEMM_WIDE": "0"})
    assert "(128 x 64)" not in narrow and "(128 x 32)" in narrow


@pytest.mark.parametrize("env", [{"OPRL_B200_GEMM_WIDE": "0"}, {"OPRL_B200_GEMM_PARTITION": "0"},
                                 {"OPRL_B200_ADAM_PATCH": "0"}, {"OPRL_B200_ADAM_PATCH": "1"},
                                 {"OPRL_B200_DW0_DEFER": "0"}],
                         ids=["narrow-tiles", "no-partition", "adam-elementwise", "adam-patches-everywhere",
                              "dw0-sum-in-the-epilogue"])
def test_plan_switches_hold_the_parity_bar(env):
    child_env = dict(os.environ, **env)
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-s",
                          "-k", "fixture_parity"], env=child_env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
