"""Where the time of a chain launch goes: clock64 stamps of CTA 0 (OPRL_B200_CHAIN_PROF=1), plus the
chain-only / GEMM-only / SIMT-only graph replays of one update.   python tools/chain_prof.py [ddpg|td3]"""
import os, sys
os.environ["OPRL_B200_CHAIN_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, make_algo, fill_buffer
from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

name = sys.argv[1] if len(sys.argv) > 1 else "ddpg"
wl = WORKLOADS[name]
S, A, B = wl["S"], wl["A"], wl["B"]
algo = make_algo(name, S, A, "cuda:0")
buf = EpisodicReplayBuffer(buffer_size_transitions=100_000, state_dim=S, action_dim=A, device="cuda:0").create()
fill_buffer(buf, 50, 0)
algo.attach_buffer(buf)
eng = algo.engine
eng.set_prefix(buf.ep_lens[:buf.episodes_counter])
for _ in range(20):
    algo.learner_step(B)
torch.cuda.synchronize()
for which in (0, 1):
    try:
        p = eng.chain_prof(B, which)
    except Exception as ex:
        print("chain", which, "no profile:", ex)
        continue
    t0 = p[0]
    print(f"chain {which}: inputs staged +{p[1]-t0}, last op done +{p[2]-t0}, exit +{p[3]-t0} cycles")
    prev = p[1]
    for i in range(16):
        if p[16 + i] == 0:
            break
        print(f"  op {i:2d}: MMA start +{p[16+i]-t0:7d}  accumulators ready +{p[32+i]-t0:7d}  (since previous ready {p[32+i]-prev:6d})")
        prev = p[32 + i]
    print(f"  feeder warp 1: waited for a free slot {p[56]} cycles, store+arrive {p[57]} cycles, over {p[58]//2} of {p[58]} chunks; "
          f"first arrivals at {[int(x - t0) for x in p[48:56]]}")
    print(f"  MMA warp: waited for filled slots {p[59]}, issued {p[60]}, waited at op boundaries {p[61]} cycles")
    if which == 0:
        print("  op 4 (16 chunks), M tile 0: MMA thread [slot seen filled, MMAs + commit issued] | feeder half 0 [top, slot free, arrived]")
        for i in range(8):
            print(f"    chunk {i}: MMA {[int(p[160 + 2 * i + k] - t0) for k in range(2)]}   feeder {[int(p[96 + 3 * i + k] - t0) for k in range(3)]}")
    if which == 0:
        for i in range(3, 8):
            print(f"    chunk {i}: arrivals of the four quarter warps of feeder half 0: {[int(p[200 + 4 * i + k] - t0) for k in range(4)]}")
    names = ["accumulators in registers", "layer epilogue done", "head partials written", "behind barrier A", "dz in operand buffer", "arrived", "finisher done", "behind barrier B"]
    print(f"  epilogue of op {p[62]} (QLOSS / DXA), cycles since its accumulators were complete:")
    base = p[32 + p[62]] if 0 <= p[62] < 16 else 0
    for k, nm in enumerate(names):
        if p[63 + k]:
            print(f"      {nm:28s} +{p[63 + k] - base}")
    print(f"  accumulator read-out done (ops 0-7): {[int(x - t0) for x in p[40:48]]}")
ms, n = eng.time_chain_only(B)
print(f"chain launches: {n} per update, {ms*1e3:.1f} us per update")
ms, n = eng.time_gemm_only(B)
print(f"gemm launches: {n} per update, {ms*1e3:.1f} us per update")
print(f"simt launches: {eng.time_simt_only(B)*1e3:.1f} us per update")
