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
