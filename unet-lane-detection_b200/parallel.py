"""Multi-GPU plumbing for the inference path: frames shard by batch, one process per GPU, no data-path
collective. torch.distributed is used only to gather results / reduce timings."""
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n frames owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_predict(predict_fn, frames: torch.Tensor, group=None) -> torch.Tensor:
    """Run predict_fn (frames[lo:hi] -> uint8 masks [n,H,W]) on this rank's shard and all-gather the masks so
    every rank returns the full [N,H,W] result in input order. Works with NCCL (GPU) and gloo (CPU tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = frames.shape[0]
    lo, hi = shard_range(n, rank, world)
    local = predict_fn(frames[lo:hi])
    if world == 1:
        return local
    sizes = [shard_range(n, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], dim=0)
