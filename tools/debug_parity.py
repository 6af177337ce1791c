"""Per-tensor gradient / parameter error of the CUDA engine against a golden fixture (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oprl_oracle as O
from tests.util import *
from tests.test_gpu_parity import make_algo, load_initial, engine_flat

name = sys.argv[1]
fx = load_case(name)
sub = int(fx["subsample"])
orc = oracle_from_fixture(fx)
spec = orc.spec
algo = make_algo(fx)
load_initial(algo, orc)
for i, nz in enumerate(fixture_noise(fx, 0)):
    algo.engine.set_noise(i, nz)
algo.update(*[x.cuda() for x in fixture_batch(fx, 0)])
run_fixture_updates(orc, fx, 0)
print({k: v for k, v in algo.engine.scalars().items()})
print(orc.scalars)
def report(tag, got, ref, shapes, nets):
    o = 0
    for n in range(nets):
        for shp in shapes:
            k = int(np.prod(shp))
            g, r = got[o:o+k], ref[o:o+k]
            err = np.abs(g - r)
            print(f"{tag} net{n} {str(shp):12s} max|ref|={np.abs(r).max():.3e} max err={err.max():.3e} "
                  f"rel={err.max()/ (np.abs(r).max()+1e-30):.2e} n_bad={(err > 1e-6 + 1e-3*np.abs(r)).sum()}")
            o += k
ar = algo.engine.arena
ga = np.concatenate([g.reshape(-1).numpy() for g in orc.last_actor_grads])
gc = np.concatenate([g.reshape(-1).numpy() for g in orc.last_critic_grads])
report("actor grad ", ar["actor"]["grad"].cpu().numpy(), ga, O.mlp_param_shapes(spec.actor_dims()), 1)
report("critic grad", ar["critic"]["grad"].cpu().numpy(), gc, O.mlp_param_shapes(spec.critic_dims()), spec.n_critics)
report("actor theta", ar["actor"]["theta"].cpu().numpy(), orc.flat("actor"), O.mlp_param_shapes(spec.actor_dims()), 1)
report("critic theta", ar["critic"]["theta"].cpu().numpy(), orc.flat("critic"), O.mlp_param_shapes(spec.critic_dims()), spec.n_critics)
report("critic target", ar["critic"]["target"].cpu().numpy(), orc.flat("critic_target"), O.mlp_param_shapes(spec.critic_dims()), spec.n_critics)
# worst actor-gradient elements
got = ar["actor"]["grad"].cpu().numpy()
err = np.abs(got - ga)
idx = np.argsort(-err / (np.abs(ga) + 1e-12))[:12]
shapes = O.mlp_param_shapes(spec.actor_dims())
offs = np.cumsum([0] + [int(np.prod(s)) for s in shapes])
for i in idx:
    t_ = int(np.searchsorted(offs, i, side="right") - 1)
    loc = np.unravel_index(i - offs[t_], shapes[t_])
    print(f"tensor {t_} {shapes[t_]} at {loc}: ref={ga[i]:+.4e} got={got[i]:+.4e}")
