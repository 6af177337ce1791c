"""Print one-update parameter L2 vs the golden fixtures for the current OPRL_B200_GEMM_MODE."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import *
from tests.test_gpu_parity import make_algo, load_initial, compare_to_fixture
for name in sys.argv[1:]:
    fx = load_case(name)
    orc = oracle_from_fixture(fx)
    algo = make_algo(fx)
    load_initial(algo, orc)
    for i, nz in enumerate(fixture_noise(fx, 0)):
        algo.engine.set_noise(i, nz)
    algo.update(*[x.cuda() for x in fixture_batch(fx, 0)])
    print(f"mode={os.environ.get('OPRL_B200_GEMM_MODE','0')} {name}: L2 after 1 update = {compare_to_fixture(algo, fx, 'first'):.3e}", flush=True)
