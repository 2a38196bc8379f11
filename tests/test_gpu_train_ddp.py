"""GPU x2 (skipped on a single-GPU box): data-parallel FusedTrainStep over NCCL. With the SAME batch on both ranks the
summed-and-rescaled gradient equals the single-GPU gradient, so both replicas must end up with (numerically) the
parameters a single-GPU step produces, and stay identical to each other."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import unet_lane_detection_b200 as U
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128]).cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g).cuda()
    y = (torch.rand(8, 1, 32, 32, generator=g) < 0.2).float().cuda()
    step = U.FusedTrainStep(net, lr=1e-3)
    for _ in range(3):
        losses = step.step(x, y)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    q.put((rank, float((others[0] - others[1]).abs().max()), flat.cpu(), losses.cpu()))
    dist.destroy_process_group()


def test_two_gpu_step_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    # replicas stay in lock step up to the run-to-run noise of fp32 atomics (identical all-reduced gradients are applied
    # to identical parameters; the only divergence is each replica's own forward nondeterminism in later steps)
    assert res[0][1] < 5e-3, res[0][1]
    # single-GPU reference in this process
    import unet_lane_detection_b200 as U
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128]).cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g).cuda()
    y = (torch.rand(8, 1, 32, 32, generator=g) < 0.2).float().cuda()
    step = U.FusedTrainStep(net, lr=1e-3)
    for _ in range(3):
        losses = step.step(x, y)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    assert (flat - res[0][2]).abs().max().item() < 5e-3
    assert (losses.cpu() - res[0][3]).abs().max().item() < 5e-3
