import sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import torch, bench
from oprl_b200 import _lib as L
from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer
wl = bench.WORKLOADS["ddpg"]; B = 256
algo = bench.make_algo("ddpg", wl["S"], wl["A"], "cuda:0")
buf = EpisodicReplayBuffer(buffer_size_transitions=1_000_000, state_dim=wl["S"], action_dim=wl["A"], device="cuda:0").create()
bench.fill_buffer(buf, 100, seed=0); algo.attach_buffer(buf)
eng = algo.engine; eng.set_prefix(buf.ep_lens[:buf.episodes_counter])
for _ in range(50): algo.learner_step(B)
torch.cuda.synchronize()
lib, h = eng._lib, eng._h
for name, fn in (("python learner_step", lambda: algo.learner_step(B)), ("raw ctypes oprl_step", lambda: lib.oprl_step(h, B, L.UPDATE_ACTOR))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%s: enqueue %.1f us/step, total %.1f us/step" % (name, (t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
