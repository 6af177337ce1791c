"""GPU: the two opt-in launch plans built on the batch-slice chain kernel (csrc/chain.cuh) hold the same parity bar
as the default stage path.  The switches are read once per process, so each plan runs the fixture / oracle parity
tests of tests/test_gpu_parity.py in a child interpreter.

  OPRL_B200_CHAIN=1        whole DDPG / TD3 update: chain(critic step) -> dW GEMM -> Adam -> chain(actor step) -> dW GEMM -> Adam
  OPRL_B200_CHAIN_ACTOR=1  critic step on the stage path, actor step as one chain launch
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LAUNCH_PROBE = """
import sys
sys.path.insert(0, %r)
from bench import make_algo
a = make_algo("ddpg", 24, 6, "cuda:0")
print("LAUNCHES", a.engine.launches(256, True))
""" % ROOT


@pytest.mark.parametrize("env,launches", [({"OPRL_B200_CHAIN": "1"}, 6), ({"OPRL_B200_CHAIN_ACTOR": "1"}, 11)],
                         ids=["full-chain", "actor-step-chain"])
def test_chain_plans_hold_the_parity_bar(env, launches):
    child_env = dict(os.environ, **env)
    out = subprocess.run([sys.executable, "-c", LAUNCH_PROBE], env=child_env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert f"LAUNCHES {launches}" in out.stdout, out.stdout  # (the default stage path: 15)
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-s",
                          "-k", "ddpg or td3"], env=child_env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
