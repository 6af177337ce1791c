#!/usr/bin/env python
"""bench.py -- gradient-updates/sec of the off-policy update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--algo ddpg|td3|sac|tqc] [--batch B]
    python bench.py --impl reference ...     # the reference algorithm's CPU path (oracle port)

One "step" = one gradient update: replay minibatch gather -> target-Q -> critic backward -> Adam
-> actor backward -> Adam -> Polyak (reference learner loop, distrib/policy_update_worker.py:66-68).

Printed JSON (one line, rank 0):
  value      whole-job updates/s with everything resident in HBM (device-side index draw),
             CUDA events on the launching stream, max over ranks.
  e2e        the same metric through the public API with HOST minibatch tensors: every step copies
             the pinned host batch H2D, runs algo.update(), and reads the critic loss back D2H
             (read-back pipelined one step deep; the blocking variant is reported beside it).
  roofline   tensor roofline of the grouped tcgen05 GEMM kernel: algorithmic FLOPs of one update
              / the time of the update's GEMM launches, timed live with CUDA
             events, against MEASURED_PEAKS.json (sustained bf16; the kernels run 3xTF32).
  cpu_baseline  the oracle port (same operator sequence as the reference's CPU PyTorch path)
             timed on this host's cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# workload table: BASELINE.json configs (dims from SURVEY.md section 8)
WORKLOADS = {
    "ddpg": dict(S=24, A=6, B=256, episodes=1000, name="DDPG walker-walk (S=24,A=6) batch 256, GPU-resident replay 1e6 transitions"),
    "td3": dict(S=17, A=6, B=256, episodes=1000, name="TD3 cheetah-run (S=17,A=6) twin critics batch 256, GPU-resident replay 1e6 transitions"),
    "sac": dict(S=67, A=21, B=1024, episodes=1000, name="SAC humanoid-stand (S=67,A=21) auto-alpha batch 1024, GPU-resident replay 1e6 transitions"),
    "tqc": dict(S=24, A=6, B=256, episodes=100, name="TQC walker-walk 5x25 quantiles batch 256, GPU-resident replay 1e5 transitions"),
}
# algorithmic MFLOP per update and gathered bytes per update
ALGO_MFLOP = {"ddpg": 365.4, "td3": 413.5, "sac": 2738.4, "tqc": 8564.8}
# dram__bytes_read.sum + dram__bytes_write.sum per gemm_kernel launch from the committed
# `ncu --set full` capture (profiles/r1b_ncu_gemm_kernel_summary.csv: 8.92 MB over the 12 launches of one
# DDPG update, cold L2 as ncu replays it; in the live loop the operands are L2-resident)
GEMM_DRAM_BYTES_PER_LAUNCH = {"ddpg": 743_147}
L_EP = 1000


class NullLogger:
    log_dir = "/tmp"

    def log_scalar(self, *a, **k):
        pass

    def log_scalars(self, *a, **k):
        pass


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm=6650.0, tf=1400.0, src="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------ GPU engine arm
def make_algo(name, S, A, device, world_size=1):
    from oprl_b200.algos.ddpg import DDPG
    from oprl_b200.algos.td3 import TD3

    classes = {"ddpg": DDPG, "td3": TD3}
    try:
        from oprl_b200.algos.sac import SAC
        from oprl_b200.algos.tqc import TQC

        classes.update(sac=SAC, tqc=TQC)
    except ImportError:
        pass
    kw = {}
    if name == "sac":
        kw["tune_alpha"] = True
    return classes[name](logger=NullLogger(), state_dim=S, action_dim=A, device=device, **kw).create()


def fill_buffer(buf, episodes, seed):
    """Synthetic replay content: full 1000-step episodes, state ~ N(0,1),
    action ~ U(-1,1), reward ~ U(0,1), done = 0 -- written on the device, bookkeeping as if every
    episode had been pushed through add_transition(..., episode_done=True at step 1000)."""
    g = torch.Generator(device=buf.states.device).manual_seed(seed)
    E = episodes
    buf.states[:E, :L_EP].normal_(generator=g)
    buf.actions[:E].uniform_(-1, 1, generator=g)
    buf.rewards[:E].uniform_(0, 1, generator=g)
    buf.dones[:E].zero_()
    for e in range(E):
        buf.ep_lens[e] = L_EP
    buf._number_transitions = E * L_EP
    buf._ep_pointer = E % buf._max_episodes
    buf.episodes_counter = min(E + 1, buf._max_episodes)


D2H_STATE_BYTES = 256  # sizeof(DevState): the block one scalar read-back copies


def run_engine(args):
    import torch.distributed as dist

    from oprl_b200.buffers.episodic_buffer import EpisodicReplayBuffer

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    wl = WORKLOADS[args.algo]
    S, A = wl["S"], wl["A"]
    B = args.batch or wl["B"]
    algo = make_algo(args.algo, S, A, device)
    buf = EpisodicReplayBuffer(buffer_size_transitions=1_000_000, state_dim=S, action_dim=A, device=device).create()
    fill_buffer(buf, wl["episodes"], seed=0 if args.mode == "dp" else rank)  # dp: replicated content
    algo.attach_buffer(buf)
    eng = algo.engine
    eng.set_prefix(buf.ep_lens[:buf.episodes_counter])
    td3 = args.algo == "td3"
    dp = world > 1 and args.mode == "dp"
    if dp:
        algo.enable_data_parallel()
    stream = torch.cuda.Stream(device=device)

    def learner_steps(n, k0=0):
        for k in range(n):
            algo.learner_step(B)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    with torch.cuda.stream(stream):
        learner_steps(args.warmup)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        t_host0 = time.perf_counter()
        learner_steps(args.steps)
        host_enqueue_us = (time.perf_counter() - t_host0) / args.steps * 1e6  # host time to ENQUEUE one step
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.result()

        # ---- e2e: host-resident pinned minibatch -> update() -> loss read-back, every step
        g = torch.Generator().manual_seed(1234)
        host = [torch.randn(B, S, generator=g), torch.rand(B, A, generator=g) * 2 - 1, torch.rand(B, 1, generator=g),
                torch.zeros(B, 1), torch.randn(B, S, generator=g)]
        host = [x.pin_memory() for x in host]
        h2d = sum(x.numel() * 4 for x in host)
        e2e_steps = max(50, min(args.steps, 1000))
        loss = 0.0

        def e2e_loop(n, depth):
            # depth 0: read the loss of update t before launching update t+1 (host stalls every step);
            # depth 1: the D2H read of update t is enqueued right behind it and consumed after
            # update t+1 has been launched -- every step still does its H2D and its D2H.
            nonlocal loss
            pending = []
            for _ in range(n):
                algo.update(*host)
                pending.append(eng.scalars_async())
                if len(pending) > depth:
                    loss = pending.pop(0).result()["critic_loss"]
            while pending:
                loss = pending.pop(0).result()["critic_loss"]

        e2e_res = {}
        for depth in (0, 1):
            e2e_loop(max(3, args.warmup // 4), depth)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            e2e_loop(e2e_steps, depth)
            e1.record(stream)
            barrier()
            e2e_res[depth] = e0.elapsed_time(e1)
        e2e_sync_ms, e2e_ms = e2e_res[0], e2e_res[1]

        # ---- API loop (GPU-resident buffer, host index draw as the reference): sample(); update()
        np.random.seed(0)
        api_steps = max(50, min(args.steps, 1000))
        for _ in range(10):
            algo.update(*buf.sample(B))
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(api_steps):
            algo.update(*buf.sample(B))
        a1.record(stream)
        barrier()
        api_ms = a0.elapsed_time(a1)

        # ---- independent learners sharing this GPU (the reference's run_training(seeds=N) mode,
        # runners/train.py:35-49): R engines, one stream each, driven round-robin from this thread.
        # One update keeps at most ~48 of the 148 SMs busy, so several seeds overlap almost freely.
        multi = None
        if args.replicas > 1 and not dp:
            algos = [algo] + [make_algo(args.algo, S, A, device) for _ in range(args.replicas - 1)]
            streams = [stream] + [torch.cuda.Stream(device=device) for _ in range(args.replicas - 1)]
            for a2 in algos[1:]:
                a2.attach_buffer(buf)
                a2.engine.set_prefix(buf.ep_lens[:buf.episodes_counter])

            def round_robin(n):
                for _ in range(n):
                    for a2, s2 in zip(algos, streams):
                        with torch.cuda.stream(s2):
                            a2.learner_step(B)

            round_robin(max(3, args.warmup // 2))
            barrier()
            t0 = time.perf_counter()
            m_steps = max(50, min(args.steps, 1000))
            round_robin(m_steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            multi = {"learners": args.replicas, "value": args.replicas * m_steps / dt, "unit": "updates/s (sum over independent learners on one GPU)",
                     "what": "R independent seeds, one CUDA stream each, same replay storage; wall clock around the loop + synchronize"}
            for a2 in algos[1:]:
                a2.engine.close()

        # ---- roofline of the dominant kernel: only the update's GEMM launches, replayed
        gemm_ms, gemm_launches = eng.time_gemm_only(B, iters=200)
        simt_ms = eng.time_simt_only(B, iters=200)
        gather_us = eng.time_gather_only(B, iters=200)

    t_ms = torch.tensor([ms, e2e_ms, api_ms, e2e_sync_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, api_ms, e2e_sync_ms = [float(x) for x in t_ms.cpu()]
    launches_per_update = eng.launches(B, True) + 1  # + gather
    if td3:
        launches_per_update = (eng.launches(B, True) + eng.launches(B, False)) / 2 + 1
    pk = peaks()
    flops = ALGO_MFLOP[args.algo] * 1e6
    achieved_tf = flops / (gemm_ms * 1e-3) / 1e12
    gather_bytes = B * (2 * S + A + 2) * 4
    out = {
        "metric": "gradient-updates/sec (batch=%d)" % B,
        "value": world * args.steps / (ms * 1e-3),
        "unit": "updates/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32 (3xTF32 split on tcgen05 kind::tf32, fp32 accumulate in TMEM)",
        "data": "synthetic",
        "config": {"workload": wl["name"], "algo": args.algo, "batch": B,
                   "parallelism": "1 learner" if world == 1 else (
                       f"dp{world}: replicated buffer + parameters, {B} rows/GPU (global minibatch {B * world}), gradient arenas "
                       f"all-reduced inside the Adam kernels over NVLink peer memory (OPRL_B200_DP_NCCL=1: NCCL all-reduce between graph "
                       f"segments); value counts {B}-row minibatch updates job-wide (optimizer steps/s = value / {world})" if dp else f"{world} independent learner replicas (one seed per GPU, no collective)"),
                   "l2": "replay storage (128 MB) exceeds L2 and is sampled uniformly; parameters/activations (~3 MB) are L2-resident by construction of the learner loop, as in the reference loop",
                   "index_draw": "device Philox (value) / host numpy (api_loop)"},
        "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": "updates/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": D2H_STATE_BYTES, "steps": e2e_steps,
                "last_critic_loss": loss,
                "what": "algo.update(*pinned_host_batch); engine.scalars_async() every step; the read-back of update t "
                        "is consumed after update t+1 was launched (one step deep)",
                "blocking_read_every_step": world * e2e_steps / (e2e_sync_ms * 1e-3)},
        "api_loop": {"value": world * api_steps / (api_ms * 1e-3), "unit": "updates/s",
                     "what": "buffer.sample(B) (host index draw, 2 KB H2D) ; algo.update(*batch) -- no per-step sync"},
        "host_enqueue_us_per_step": host_enqueue_us,
        "gpu_launches": int(round(launches_per_update * args.steps)),
        "launches_per_update": launches_per_update,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": pk["tf"], "unit": "TFLOP/s",
                     "frac": achieved_tf / pk["tf"], "traffic": GEMM_DRAM_BYTES_PER_LAUNCH.get(args.algo),
                     "kernel": "oprl::gemm_kernel<false> (grouped 128x32 tcgen05 3xTF32 tiles)",
                     "launches_per_update": gemm_launches, "gemm_us_per_update": gemm_ms * 1e3,
                     "simt_us_per_update": simt_ms * 1e3,
                     "algorithmic_mflop_per_update": ALGO_MFLOP[args.algo], "peak_source": pk["src"],
                     "achieved_per_launch_note": "achieved = algorithmic FLOPs of one update / summed duration of its GEMM launches (CUDA events around a GEMM-only graph replay); traffic = DRAM bytes per launch under ncu (cold L2)",
                     "note": "latency-bound by construction: %.1f MFLOP/update is %.2f us of tensor time at the measured peak; tf32 peak is 1/2 of bf16 and 3xTF32 needs 3 passes" % (ALGO_MFLOP[args.algo], ALGO_MFLOP[args.algo] * 1e6 / (pk["tf"] * 1e12) * 1e6)},
        "roofline_gather": {"bound": "hbm", "achieved": gather_bytes / (gather_us * 1e-6) / 1e9, "peak": pk["hbm"],
                            "unit": "GB/s", "frac": gather_bytes / (gather_us * 1e-6) / 1e9 / pk["hbm"],
                            "bytes_per_launch": gather_bytes, "us_per_launch": gather_us},
        "clocks": clocks,
    }
    if multi:
        out["multi_learner"] = multi
    if rank == 0:
        if args.cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.algo, B, budget_s=12.0)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------- CPU reference arm
def oracle_learner(algo, B, episodes=20):
    """The reference learner loop on the CPU oracle port: sample(B) ; update(*batch)."""
    from oracle import oprl_oracle as O

    wl = WORKLOADS[algo]
    S, A = wl["S"], wl["A"]
    spec = O.AlgoSpec(algo=algo, state_dim=S, action_dim=A, tune_alpha=(algo in ("sac", "tqc")),
                      lr_alpha=3e-4 if algo == "tqc" else 1e-3)
    actor, critics = O.init_params(spec, 0)
    orc = O.OracleAlgo(spec, actor, critics)
    st, ac, rw, dn = [torch.from_numpy(x) for x in O.synthetic_buffer(episodes, L_EP, S, A, 0)]
    ep_lens = [L_EP] * episodes
    n_noise = {"ddpg": 0, "td3": 1, "sac": 2, "tqc": 2}[algo]

    def step():
        inds = np.random.randint(0, episodes * L_EP, size=B)
        ep, sp = O.inds_to_episodic(inds, ep_lens, episodes)
        batch = (st[ep, sp], ac[ep, sp], rw[ep, sp], dn[ep, sp], st[ep, sp + 1])
        noise = [torch.randn(B, A) for _ in range(n_noise)]
        orc.update(*batch, noise=noise)

    return step


def cpu_baseline(algo, B, budget_s):
    """All host threads (the headline figure) and, beside it, one thread (SURVEY.md section 8d asks for both)."""
    step = oracle_learner(algo, B)
    out = {}
    for threads, budget in ((os.cpu_count() or 1, budget_s), (1, budget_s / 3)):
        torch.set_num_threads(threads)
        for _ in range(5):
            step()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget:
            step()
            n += 1
        dt = time.perf_counter() - t0
        if not out:
            out = {"value": n / dt, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n} updates of the same workload (20x1000-step replay, batch {B}) in {dt:.1f} s, torch {torch.__version__} CPU"}
        else:
            out["value_1_thread"] = n / dt
    torch.set_num_threads(os.cpu_count() or 1)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.algo]
    B = args.batch or wl["B"]
    torch.set_num_threads(os.cpu_count() or 1)
    np.random.seed(0)
    torch.manual_seed(0)
    step = oracle_learner(args.algo, B)
    # bounded: keep the whole run within a few minutes whatever K is asked for
    warm = min(args.warmup, 20)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    probe = 5
    for _ in range(probe):
        step()
    per = (time.perf_counter() - t0) / probe
    steps = max(10, min(args.steps, int(120.0 / per)))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = steps / dt
    cores = torch.get_num_threads()
    sample = f"{steps} updates (asked {args.steps}) of sample(B);update on the CPU oracle port, batch {B}, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "gradient-updates/sec (batch=%d)" % B, "value": val, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "algo": args.algo, "batch": B},
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--algo", default="ddpg", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--replicas", type=int, default=1,
                    help="also time R independent learners (seeds) sharing each GPU; reported separately")
    ap.add_argument("--mode", default="dp", choices=["dp", "replicas"],
                    help="N>1: data-parallel learners with gradient all-reduce (default) or independent replicas")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
