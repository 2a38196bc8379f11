"""CPU, world_size 2 over gloo: the batch-sharding plumbing used by the multi-GPU inference path."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet_lane_detection_b200.parallel import sharded_predict
    frames = torch.arange(n * 4 * 4 * 3, dtype=torch.int64).reshape(n, 4, 4, 3).to(torch.uint8)

    def fake_predict(x):  # stand-in for the GPU path: any per-frame function must come back in input order
        return ((x.sum(dim=3) % 2) * 255).to(torch.uint8)

    out = sharded_predict(fake_predict, frames)
    ok = torch.equal(out, fake_predict(frames))
    q.put((rank, bool(ok), tuple(out.shape)))
    dist.destroy_process_group()


def _run(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape[0] == n for _, _, shape in res)


def test_sharded_predict_even_split():
    _run(8)


def test_sharded_predict_ragged_and_tiny():
    _run(7)
    _run(1)


def _grad_worker(rank, world, port, q):
    """Data-parallel training exchange: all-reduce(sum) of the flat gradient + 1/world folded into AdamW
    must equal single-process AdamW on the mean gradient."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet_lane_detection_b200.training import allreduce_gradients
    gen = torch.Generator().manual_seed(100)
    p0 = torch.randn(1000, generator=gen)
    grads = [torch.randn(1000, generator=gen) for _ in range(world)]      # every rank can rebuild every replica's gradient
    mine = grads[rank].clone()
    scale = allreduce_gradients(mine)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-4, weight_decay=1e-4)
    ref.grad = torch.stack(grads).mean(0)
    opt.step()
    # python restatement of the AdamW kernel's arithmetic (csrc/train_kernels.cuh adamw_kernel), step 1
    g = mine * scale
    p = p0 * (1 - 1e-4 * 1e-4)
    m, v = 0.1 * g, 0.001 * g * g
    p = p - (1e-4 / (1 - 0.9)) * (m / (v.sqrt() / (1 - 0.999) ** 0.5 + 1e-8))
    q.put((rank, float((p - ref.detach()).abs().max()), scale))
    dist.destroy_process_group()


def test_gradient_allreduce_folds_world_size_into_adamw():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(err < 1e-6 and scale == 0.5 for _, err, scale in res), res
