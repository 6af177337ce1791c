#!/usr/bin/env python
"""bench.py -- gradient-updates/sec of the off-policy update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--algo ddpg|td3|sac|tqc] [--batch B]
    python bench.py --impl reference ...     # the reference's own CPU path (baseline/_ref), else the oracle port

One "step" = one gradient update: replay minibatch gather -> target-Q -> critic backward -> Adam
-> actor backward -> Adam -> Polyak (reference learner loop, distrib/policy_update_worker.py:66-68).

Printed JSON (one line, rank 0):
  value      whole-job updates/s with everything resident in HBM (device-side index draw),
             CUDA events on the launching stream, max over ranks.
  e2e        the same metric through the public API with HOST minibatch tensors: every step copies
             the pinned host batch H2D, runs algo.update(), and reads the critic loss back D2H
             (read-back pipelined one step deep; the blocking variant is reported beside it).
  roofline   tensor roofline of the dominant kernel (the batch-slice chain kernel for DDPG / TD3, the
             grouped tcgen05 GEMM kernel otherwise): algorithmic FLOPs that kernel's launches perform
             in one update / their duration, timed live with CUDA events around a replay of just those
             launches, against MEASURED_PEAKS.json (sustained bf16; the kernels run 3xTF32).
  cpu_baseline  the reference itself (unmodified, installed into baseline/_ref by __graft_entry__.build();
             kind "reference") -- or the oracle port when that install is missing (kind "port") -- timed
             on this host's cores on a bounded sample of the same workload: best of {all threads, 1 thread}.
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
L_EP = 1000
MIN_SETTLE_STEPS = 16  # untimed steps before the timed window whatever --warmup says (graph captures, lazy allocations)


def ncu_traffic_per_launch(kernel_substr, algo=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the newest committed
    `ncu --set full` raw-page CSV under profiles/ (cold L2: ncu flushes caches between replays).  None if absent.
    The raw page is: a header row, a units row, then one row per profiled launch."""
    import csv
    import glob

    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0}
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*raw*.csv")), reverse=True)
    tag = lambda p_: next((a for a in WORKLOADS if f"_{a}_" in os.path.basename(p_)), None)
    # a capture of this algorithm's own update first (file name carries _<algo>_), then the untagged ones (DDPG)
    paths = [p_ for p_ in paths if algo and tag(p_) == algo] + [p_ for p_ in paths if tag(p_) is None]
    for path in paths:
        try:
            rows = [r for r in csv.reader(open(path)) if r]
            hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
            cols = {name: k for k, name in enumerate(rows[hdr])}
            units = rows[hdr + 1] if hdr + 1 < len(rows) else []
            kn, rd, wr = cols["Kernel Name"], cols["dram__bytes_read.sum"], cols["dram__bytes_write.sum"]
        except (StopIteration, KeyError, OSError):
            continue
        unit = lambda c: scale.get(units[c].strip() if c < len(units) else "", 1.0)
        tot, n = 0.0, 0
        for r in rows[hdr + 2:]:
            if len(r) <= max(kn, rd, wr) or kernel_substr not in r[kn]:
                continue
            try:
                tot += float(r[rd].replace(",", "")) * unit(rd) + float(r[wr].replace(",", "")) * unit(wr)
                n += 1
            except ValueError:
                continue
        if n:
            return {"bytes_per_launch": tot / n, "launches": n, "source": os.path.relpath(path, ROOT)}
    return None


def workload_config(args, world, dp):
    """The `config` object -- identical for the engine arm and the reference arm."""
    wl = WORKLOADS[args.algo]
    B = args.batch or wl["B"]
    return {"workload": wl["name"], "algo": args.algo, "batch": B, "replay_transitions": wl["episodes"] * L_EP,
            "parallelism": "1 learner" if world == 1 else (f"dp{world}: {B} rows per learner, global minibatch {B * world}" if dp else f"{world} independent learner replicas"),
            "l2": "replay storage exceeds L2 and is sampled uniformly; parameters / activations are L2-resident by construction of the learner loop, as in the reference loop"}


def mac_counts(algo, S, A):
    """(forward + dX MACs, dW MACs) per batch row of one full update, from the layer shapes."""
    H = 256
    actor = [(S, H), (H, H), (H, A)]
    critic = [(S + A, H), (H, H), (H, 1)]
    mac = lambda net: sum(i * o for i, o in net)
    nq = {"ddpg": 1, "td3": 2}.get(algo)
    if nq is None:
        return None
    fwd = 2 * mac(actor) + (2 * nq + 1) * mac(critic)
    dx = nq * (mac(critic[1:])) + mac(critic) + mac(actor[1:])
    dw = nq * mac(critic) + mac(actor)
    if algo == "td3":  # the actor step runs every second update (td3.py:79)
        half = (mac(actor) + mac(critic)) + (mac(critic) + mac(actor[1:]))
        fwd_dx = fwd + dx - half / 2
        return fwd_dx, nq * mac(critic) + mac(actor) / 2
    return fwd + dx, dw


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

    dp_parity = None
    if dp:
        with torch.cuda.stream(stream):
            dp_parity = dp_parity_check(args.algo, device, rank, world)
    settle = max(args.warmup, MIN_SETTLE_STEPS)
    with torch.cuda.stream(stream):
        learner_steps(settle)
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
            e2e_loop(max(MIN_SETTLE_STEPS, args.warmup // 4), depth)
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
        for _ in range(MIN_SETTLE_STEPS):
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

            round_robin(max(MIN_SETTLE_STEPS, args.warmup // 2))
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

        # ---- roofline of the dominant kernel: only that kernel's launches of one update, replayed
        gemm_ms, gemm_launches = eng.time_gemm_only(B, iters=200)
        chain_ms, chain_launches = eng.time_chain_only(B, iters=200)
        simt_ms = eng.time_simt_only(B, iters=200)
        gather_us = eng.time_gather_only(B, iters=200)

    # ---- N > 1: the two other scaling modes SURVEY.md section 8e asks for, beside the weak-scaling headline
    other_modes = None
    if dp:
        other_modes = {}
        with torch.cuda.stream(stream):
            sub_steps = max(50, min(args.steps, 500))
            # strong scaling: the reference's global minibatch B split over the ranks (B / world rows each)
            if B % world == 0 and B // world >= 1:
                Bs = B // world
                for _ in range(MIN_SETTLE_STEPS):
                    algo.learner_step(Bs)
                barrier()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(stream)
                for _ in range(sub_steps):
                    algo.learner_step(Bs)
                s1.record(stream)
                barrier()
                t_ = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                other_modes["strong"] = {"value": sub_steps / (float(t_) * 1e-3), "unit": f"optimizer steps/s on a global minibatch of {B} ({Bs} rows per GPU)",
                                         "ms_per_step": float(t_) / sub_steps, "steps": sub_steps}
            # replicas: one independent learner per GPU, no collective (the reference's run_training(seeds=N), train.py:35-49)
            solo = make_algo(args.algo, S, A, device)
            solo.attach_buffer(buf)
            solo.engine.set_prefix(buf.ep_lens[:buf.episodes_counter])
            for _ in range(MIN_SETTLE_STEPS):
                solo.learner_step(B)
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(sub_steps):
                solo.learner_step(B)
            r1.record(stream)
            barrier()
            t_ = torch.tensor([r0.elapsed_time(r1)], device=device, dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            other_modes["replicas"] = {"value": world * sub_steps / (float(t_) * 1e-3), "unit": f"updates/s summed over {world} independent learners (batch {B} each, no collective)",
                                       "ms_per_step": float(t_) / sub_steps, "steps": sub_steps}
            solo.engine.close()
            algo.attach_buffer(buf)  # (the buffer's fused sample() goes back to the data-parallel learner's engine)

    t_ms = torch.tensor([ms, e2e_ms, api_ms, e2e_sync_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, api_ms, e2e_sync_ms = [float(x) for x in t_ms.cpu()]
    launches_per_update = eng.launches(B, True) + 1  # + gather
    if td3:
        launches_per_update = (eng.launches(B, True) + eng.launches(B, False)) / 2 + 1
    pk = peaks()
    gather_bytes = B * (2 * S + A + 2) * 4
    note = "latency-bound by construction: %.1f MFLOP/update is %.2f us of tensor time at the measured peak; tf32 peak is 1/2 of bf16 and 3xTF32 needs 3 passes" % (
        ALGO_MFLOP[args.algo], ALGO_MFLOP[args.algo] * 1e6 / (pk["tf"] * 1e12) * 1e6)

    def tensor_roofline(kernel, substr, mflop, k_ms, k_launches):
        tf = mflop * 1e6 / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        tr = ncu_traffic_per_launch(substr, args.algo)
        return {"bound": "tensor", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "kernel": kernel, "launches_per_update": k_launches, "us_per_update": k_ms * 1e3,
                "us_per_launch": k_ms * 1e3 / max(k_launches, 1), "algorithmic_mflop_per_update_in_this_kernel": mflop,
                "peak_source": pk["src"],
                "achieved_per_launch_note": "achieved = algorithmic FLOPs this kernel's launches perform in one update / their summed duration "
                                            "(CUDA events around a graph replay of just those launches); traffic = DRAM bytes per launch under ncu (cold L2)",
                "note": note}

    macs = mac_counts(args.algo, S, A)
    if chain_launches > 0 and macs:
        chain_mflop, dw_mflop = 2 * B * macs[0] / 1e6, 2 * B * macs[1] / 1e6
        roof = tensor_roofline("oprl::chain_kernel (batch-slice layer chain: weights = MMA M side through TMEM, 16 batch rows = N; tcgen05 3xTF32)",
                               "chain_kernel", chain_mflop, chain_ms, chain_launches)
        roof_gemm = tensor_roofline("oprl::gemm_kernel<false> (grouped 128x32 tcgen05 3xTF32 tiles: the weight-gradient products)",
                                    "gemm_kernel", dw_mflop, gemm_ms, gemm_launches)
    else:
        tiles = "128x32 tiles" if args.algo in ("ddpg", "td3") else "128x32 tiles, 128x64 in launches that exceed one wave of SMs"
        roof = tensor_roofline("oprl::gemm_kernel<false, 1|2> (grouped tcgen05 3xTF32 GEMM, %s)" % tiles, "gemm_kernel",
                               ALGO_MFLOP[args.algo], gemm_ms, gemm_launches)
        roof_gemm = None
    roof["simt_us_per_update"] = simt_ms * 1e3
    roof["gemm_us_per_update"] = gemm_ms * 1e3
    roof["chain_us_per_update"] = chain_ms * 1e3
    roof["algorithmic_mflop_per_update"] = ALGO_MFLOP[args.algo]
    value = world * args.steps / (ms * 1e-3)
    e2e_value = world * e2e_steps / (e2e_ms * 1e-3)
    cfg = workload_config(args, world, dp)
    out = {
        "metric": "gradient-updates/sec (batch=%d)" % B,
        "value": value,
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
        "config": cfg,
        "engine_notes": {"settle_steps_before_timing": settle,
                         "index_draw": "device Philox (value) / host numpy (api_loop)",
                         "dp": (f"replicated buffer + parameters, gradient arenas all-reduced inside the Adam kernels over NVLink peer memory "
                                f"(OPRL_B200_DP_NCCL=1: NCCL all-reduce between graph segments); value counts {B}-row minibatch updates job-wide "
                                f"(optimizer steps/s = value / {world})") if dp else None},
        "e2e": {"value": e2e_value, "unit": "updates/s",
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
        "roofline": roof,
        "roofline_gather": {"bound": "hbm", "achieved": gather_bytes / (gather_us * 1e-6) / 1e9, "peak": pk["hbm"],
                            "unit": "GB/s", "frac": gather_bytes / (gather_us * 1e-6) / 1e9 / pk["hbm"],
                            "bytes_per_launch": gather_bytes, "us_per_launch": gather_us},
        "self_check": {"e2e_le_value_x1.02": bool(e2e_value <= value * 1.02),
                       "note": "e2e does strictly more work per step than value (H2D of the batch + scalar read-back)"},
        "clocks": clocks,
    }
    if roof_gemm:
        out["roofline_gemm"] = roof_gemm
    if dp_parity is not None:
        out["dp_parity"] = dp_parity
    if other_modes:
        out["scaling_modes"] = {"weak (headline `value`)": f"{B} rows per GPU, global minibatch {B * world}", **other_modes}
    if multi:
        out["multi_learner"] = multi
    if rank == 0:
        if args.cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.algo, B, budget_s=16.0)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------- data-parallel parity record
def dp_parity_check(algo_name, device, rank, world):
    """One update of the reference-generated golden fixture with its rows split over the ranks, through the
    fused NVLink path and through the NCCL path: both must reproduce the reference's FULL-batch update
    (batch means of ddpg.py:98,104).  Returns {path: {l2, loss_err}}; the run fails if l2 > 1e-5."""
    import torch.distributed as dist

    from tests.test_gpu_parity import compare_to_fixture, load_initial, make_algo as make_fx_algo
    from tests.util import fixture_batch, fixture_noise, load_case, oracle_from_fixture

    name = algo_name if algo_name in ("ddpg", "td3") else "ddpg"
    fx = load_case(name)
    out = {"fixture": f"tests/golden/{name}.npz (first update, {fx['s0'].shape[0]} rows split over {world} ranks)"}
    for path, fused in (("fused_nvlink_adam", True), ("nccl_allreduce", False)):
        orc = oracle_from_fixture(fx)
        a = make_fx_algo(fx, device=device)
        load_initial(a, orc)
        a.enable_data_parallel(fused=fused)
        batch = fixture_batch(fx, 0)
        Bf = batch[0].shape[0]
        lo, hi = rank * Bf // world, (rank + 1) * Bf // world
        for i, nz in enumerate(fixture_noise(fx, 0)):
            a.engine.set_noise(i, nz[lo:hi])
        a.update(*[x[lo:hi].to(device) for x in batch])
        l2 = float(compare_to_fixture(a, fx, "first"))
        sc = a.engine.scalars()
        t_ = torch.tensor([sc["critic_loss"], sc["actor_loss"], l2], device=device, dtype=torch.float64)
        mx = t_.clone()
        dist.all_reduce(t_)  # per-rank loss scalars are shares of the global means
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        loss_err = max(abs(float(t_[0]) - float(fx["scalar0_critic_loss"])), abs(float(t_[1]) - float(fx["scalar0_actor_loss"])))
        out[path] = {"l2": float(mx[2]), "loss_err": loss_err}
        a.engine.close()
        dist.barrier()
        if float(mx[2]) > 1e-5 or loss_err > 1e-4:
            raise SystemExit(f"data-parallel parity FAILED on the {path} path: l2 {float(mx[2]):.3e}, loss err {loss_err:.3e}")
    out["l2"] = max(out["fused_nvlink_adam"]["l2"], out["nccl_allreduce"]["l2"])
    out["loss_err"] = max(out["fused_nvlink_adam"]["loss_err"], out["nccl_allreduce"]["loss_err"])
    out["path"] = "fused_nvlink_adam (timed) + nccl_allreduce (baseline)"
    return out


# ------------------------------------------------------------------- CPU reference arm
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_learner(algo, B):
    """The reference's own learner loop body (distrib/policy_update_worker.py:66-68) on its own classes:
    `batch = buffer.sample(B); algo.update(*batch)` -- UNMODIFIED reference code imported from baseline/_ref
    (installed by __graft_entry__.build() from /root/reference), device cpu, same replay size as the GPU arm."""
    if not os.path.isdir(os.path.join(REF_DIR, "oprl")):
        raise ImportError("baseline/_ref holds no reference install")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from oprl.algos.ddpg import DDPG
    from oprl.algos.sac import SAC
    from oprl.algos.td3 import TD3
    from oprl.algos.tqc import TQC
    from oprl.buffers.episodic_buffer import EpisodicReplayBuffer as RefBuffer

    wl = WORKLOADS[algo]
    S, A, E = wl["S"], wl["A"], wl["episodes"]
    kw = {"tune_alpha": True} if algo == "sac" else {}
    ref = dict(ddpg=DDPG, td3=TD3, sac=SAC, tqc=TQC)[algo](logger=NullLogger(), state_dim=S, action_dim=A, device="cpu", **kw).create()
    buf = RefBuffer(buffer_size_transitions=1_000_000, state_dim=S, action_dim=A, device="cpu").create()
    g = torch.Generator().manual_seed(0)
    buf.states[:E, :L_EP].normal_(generator=g)
    buf.states[:E, L_EP].zero_()
    buf.actions[:E].uniform_(-1, 1, generator=g)
    buf.rewards[:E].uniform_(0, 1, generator=g)
    buf.dones[:E].zero_()
    for ep in range(E):
        buf.ep_lens[ep] = L_EP
    buf._number_transitions = E * L_EP
    buf._ep_pointer = E % buf._max_episodes
    buf.episodes_counter = min(E + 1, buf._max_episodes)

    def step():
        ref.update(*buf.sample(B))

    return step


def oracle_learner(algo, B):
    """Fallback when the reference install is absent: the same loop on the CPU oracle port."""
    from oracle import oprl_oracle as O

    wl = WORKLOADS[algo]
    S, A, episodes = wl["S"], wl["A"], wl["episodes"]
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


def cpu_learner(algo, B):
    """(step function, kind): the reference itself when baseline/_ref is there, else the oracle port."""
    try:
        return reference_learner(algo, B), "reference"
    except Exception as ex:  # missing install / missing third-party import of the reference package
        sys.stderr.write(f"bench: reference install unusable ({type(ex).__name__}: {ex}); timing the oracle port\n")
        return oracle_learner(algo, B), "port"


def time_cpu(step, threads, budget_s, warm=5):
    torch.set_num_threads(threads)
    for _ in range(warm):
        step()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        step()
        n += 1
    return n, time.perf_counter() - t0


def cpu_baseline(algo, B, budget_s):
    """Best of {all host threads, one thread} (on these small shapes MKL oversubscription makes all threads
    the slower one on many hosts); both figures are reported."""
    step, kind = cpu_learner(algo, B)
    wl = WORKLOADS[algo]
    n_all = os.cpu_count() or 1
    res = {}
    for threads in (n_all, 1):
        n, dt = time_cpu(step, threads, budget_s / 2)
        res[threads] = (n / dt, n, dt)
    torch.set_num_threads(n_all)
    best = max(res, key=lambda k: res[k][0])
    v, n, dt = res[best]
    return {"value": v, "unit": "updates/s", "cores": best, "kind": kind,
            "value_all_threads": res[n_all][0], "threads_all": n_all, "value_1_thread": res[1][0],
            "sample": f"{n} updates of `buffer.sample({B}); algo.update(*batch)` on a {wl['episodes'] * L_EP}-transition CPU replay in {dt:.1f} s "
                      f"({best} thread(s), the faster of {n_all} and 1), torch {torch.__version__} CPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.algo]
    B = args.batch or wl["B"]
    np.random.seed(0)
    torch.manual_seed(0)
    step, kind = cpu_learner(args.algo, B)
    n_all = os.cpu_count() or 1
    # pick the faster thread count on a short probe (all threads vs one), then time the bounded run with it
    probe = {}
    for threads in (n_all, 1):
        n, dt = time_cpu(step, threads, 4.0, warm=min(args.warmup, 20))
        probe[threads] = n / dt
    cores = max(probe, key=probe.get)
    torch.set_num_threads(cores)
    per = 1.0 / probe[cores]
    steps = max(10, min(args.steps, int(100.0 / per)))  # bounded: the whole run stays within a few minutes
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = steps / dt
    world = int(os.environ.get("WORLD_SIZE", 1))
    sample = (f"{steps} updates (asked {args.steps}) of `buffer.sample({B}); algo.update(*batch)` on a {wl['episodes'] * L_EP}-transition CPU replay, "
              f"{cores} thread(s) (probe: {probe[n_all]:.0f} updates/s at {n_all} threads, {probe[1]:.0f} at 1)")
    print(json.dumps({
        "impl": "reference", "metric": "gradient-updates/sec (batch=%d)" % B, "value": val, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 20), "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, world > 1 and args.mode == "dp"),
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": kind, "sample": sample,
                         "value_all_threads_probe": probe[n_all], "value_1_thread_probe": probe[1]},
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
