#!/usr/bin/env python
"""In-situ cost of every stage of one update: time the first k stages of the update graph for
k = 1..n (CUDA events around 300 replays each); the difference of consecutive prefixes is what stage k
adds to the chain, launch gap and programmatic-dependent-launch overlap included.

    python tools/stage_profile.py [--algo ddpg] [--batch 256]

The prefixes run real kernels on a real batch but leave the engine in a meaningless state: profiling
only, in its own process."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from oprl_b200 import _lib as L  # noqa: E402
from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="ddpg")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.algo]
    B = args.batch or wl["B"]
    algo = bench.make_algo(args.algo, wl["S"], wl["A"], "cuda:0")
    buf = EpisodicReplayBuffer(buffer_size_transitions=1_000_000, state_dim=wl["S"], action_dim=wl["A"], device="cuda:0").create()
    bench.fill_buffer(buf, 100, seed=0)
    algo.attach_buffer(buf)
    eng = algo.engine
    eng.set_prefix(buf.ep_lens[:buf.episodes_counter])
    for _ in range(20):
        algo.learner_step(B)
    torch.cuda.synchronize()
    eng._use_current_stream()
    prev = 0.0
    k = 1
    print(f"{args.algo} batch {B}: stage prefix timings (us per replay)")
    while True:
        ms, n = C.c_float(), C.c_int()
        L.check(eng._lib.oprl_profile(eng._h, B, L.UPDATE_ACTOR, 100 + k, args.iters, C.byref(ms), C.byref(n)))
        us = ms.value / args.iters * 1e3
        print(f"  stages 1..{k:2d}: {n.value:2d} launches {us:8.2f} us   (+{us - prev:6.2f})")
        if k > 1 and n.value == last_n:
            break
        last_n, prev = n.value, us
        k += 1
        if k > 64:
            break


if __name__ == "__main__":
    main()
