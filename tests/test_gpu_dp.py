"""GPU (>= 2 devices): two data-parallel learners, each on half of a golden minibatch (DDPG; TQC, whose large critics
take the Adam kernel's 32 x 32 patch path and the wide GEMM tiles), gradients all-reduced inside the Adam kernel over
NVLink or by NCCL -- must reproduce the reference's full-batch update."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.util import fixture_batch, fixture_noise, load_case, oracle_from_fixture

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, fused, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from tests.test_gpu_parity import compare_to_fixture, compare_to_fixture_full, load_initial, make_algo

    fx = load_case(name)
    orc = oracle_from_fixture(fx)
    algo = make_algo(fx, device=f"cuda:{rank}")
    load_initial(algo, orc)
    algo.enable_data_parallel(fused=fused)
    assert algo.engine.fused_comm == fused
    batch = fixture_batch(fx, 0)
    B = batch[0].shape[0]
    lo, hi = rank * B // world, (rank + 1) * B // world
    for i, nz in enumerate(fixture_noise(fx, 0)):
        algo.engine.set_noise(i, nz[lo:hi])
    algo.update(*[x[lo:hi].cuda() for x in batch])
    if "firstfull_critic" in fx:
        l2 = compare_to_fixture_full(algo, fx, {"critic": orc.flat("critic").copy()})
    else:
        l2 = compare_to_fixture(algo, fx, "first")
    sc = algo.engine.scalars()
    t_ = torch.tensor([sc["critic_loss"], sc["actor_loss"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t_)  # per-rank scalars are shares of the global means
    if rank == 0:
        out.put((l2, t_.cpu().numpy()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["ddpg", "tqc"])
@pytest.mark.parametrize("fused", [True, False], ids=["fused-nvlink-adam", "nccl-allreduce"])
def test_two_learners_reproduce_the_full_batch_update(fused, name):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, fused, out)) for r in range(2)]
    for p in procs:
        p.start()
    l2, losses = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fx = load_case(name)
    print(f"dp2 {name} fused={fused}: param L2 after 1 update = {l2:.3e}, losses {losses}")
    assert l2 <= 1e-5
    assert abs(losses[0] - float(fx["scalar0_critic_loss"])) <= 1e-4
    assert abs(losses[1] - float(fx["scalar0_actor_loss"])) <= 1e-4
