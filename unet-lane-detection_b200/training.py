"""Training step on the B200 kernels (reference: README.md:2060-2084 train_one_epoch, 1855-1893 BCEDiceLoss,
2173-2174 AdamW).

Two ways in, both through libunet_b200.so (no PyTorch/cuDNN compute, no CPU fallback):

* drop-in autograd: `model.train(); out = model(images); loss = criterion(out, masks); loss.backward();
  optimizer.step()` works unchanged - `UNet.forward` in training mode runs `unet_b200_train_forward` inside a
  `torch.autograd.Function` whose backward is `unet_b200_train_backward`, so the reference's own BCEDiceLoss and
  `torch.optim.AdamW` can stay.
* fused step: `FusedTrainStep(model).step(images, masks)` additionally runs the loss (+ its gradient), the
  data-parallel gradient exchange and AdamW as hand-written kernels. With more than one rank the backward runs in stages and
  the exchange of a gradient bucket starts on a side stream as soon as the bucket is final (decoder first), so it hides under
  the rest of the backward.

Parameters are kept in ONE flat fp32 buffer in `model.parameters()` order; every nn.Parameter is a view into it, so
state_dict / optimizers see ordinary tensors.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.distributed as dist

from ._lib import check, lib


class _TrainEngine:
    """One bound trainer: (device, batch, H, W) + activation/gradient workspace."""

    def __init__(self, model, device, B, H, W):
        feats = (C.c_int * len(model.features))(*model.features)
        handle = C.c_void_p()
        check(lib.unet_b200_trainer_create(C.byref(handle), B, H, W, model.in_channels, model.out_channels, feats,
                                           len(model.features)))
        self.handle = handle
        self.B, self.H, self.W, self.device = B, H, W, device
        self.n_params = lib.unet_b200_trainer_num_params(handle)
        nt = lib.unet_b200_trainer_num_tensors(handle)
        self.offsets = [lib.unet_b200_trainer_tensor_offset(handle, i) for i in range(nt + 1)]
        self.workspace = torch.empty(lib.unet_b200_trainer_workspace_bytes(handle) + 1024, dtype=torch.uint8, device=device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        check(lib.unet_b200_trainer_bind(handle, base))
        self.fwd_seq = 0   # bumped by every train-mode forward: a backward must belong to the latest one
        self.n_stages = lib.unet_b200_trainer_num_stages(handle)
        self.stage_ranges = []
        for s in range(self.n_stages):
            lo, hi = C.c_longlong(), C.c_longlong()
            check(lib.unet_b200_trainer_stage_range(handle, s, C.byref(lo), C.byref(hi)))
            self.stage_ranges.append((int(lo.value), int(hi.value)))

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            lib.unet_b200_trainer_destroy(h)
            self.handle = None


def flatten_parameters_(model, alloc=None):
    """Make every parameter of `model` a view into one contiguous fp32 buffer (parameters() order) and return it.
    Parameter objects keep their identity, so optimizers created before or after stay valid. `alloc(n, device)` lets the
    caller place the buffer (e.g. in NVLink-shared symmetric memory); once placed, later calls keep it."""
    params = list(model.parameters())
    flat = getattr(model, "_b200_flat", None)
    off, ok = 0, flat is not None
    if ok:
        for p in params:
            if p.data_ptr() != flat.data_ptr() + off * 4 or p.dtype != torch.float32:
                ok = False
                break
            off += p.numel()
        ok = ok and off == flat.numel()
    if ok:
        return flat
    dev = params[0].device
    n_total = sum(p.numel() for p in params)
    flat = alloc(n_total, dev) if alloc is not None else torch.empty(n_total, dtype=torch.float32, device=dev)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            view = flat[off:off + n].view(p.shape)
            view.copy_(p.detach().to(torch.float32))
            p.data = view
            off += n
    model._b200_flat = flat
    return flat


MAX_ENGINES = 4   # bound plans / trainers kept per model (each owns a workspace): least recently used goes first


def _lru_get(cache: OrderedDict, key, make, limit):
    hit = cache.get(key)
    if hit is not None:
        cache.move_to_end(key)
        return hit
    while len(cache) >= limit:
        cache.popitem(last=False)
    val = make()
    cache[key] = val
    return val


def _engine(model, device, B, H, W):
    def make():
        eng = _TrainEngine(model, device, B, H, W)
        sizes = [p.numel() for p in model.parameters()]
        want = [eng.offsets[i + 1] - eng.offsets[i] for i in range(len(eng.offsets) - 1)]
        if sizes != want:
            raise RuntimeError("UNet (B200): parameter layout of the module does not match the library's trainer")
        return eng
    return _lru_get(model._engines, ("train", str(device), B, H, W), make, MAX_ENGINES)


def _bn_tables(model):
    bns = [bn for _, bn in model._double_convs()]
    n = len(bns)
    means = (C.c_void_p * n)(*[bn.running_mean.data_ptr() for bn in bns])
    vars_ = (C.c_void_p * n)(*[bn.running_var.data_ptr() for bn in bns])
    return bns, means, vars_


def train_forward(model, x4):
    """x4: bf16 NHWC4 [B,H,W,4] -> logits fp32 [B,H,W] (out_channels == 1, the reference's case) or NCHW [B,out_channels,H,W];
    updates BN running statistics like nn.BatchNorm2d."""
    B, H, W, _ = x4.shape
    eng = _engine(model, x4.device, B, H, W)
    flat = flatten_parameters_(model)
    bns, means, vars_ = _bn_tables(model)
    shape = (B, H, W) if model.out_channels == 1 else (B, model.out_channels, H, W)
    logits = torch.empty(shape, dtype=torch.float32, device=x4.device)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.unet_b200_train_forward(eng.handle, x4.data_ptr(), flat.data_ptr(), means, vars_, float(bns[0].momentum),
                                      float(bns[0].eps), logits.data_ptr(), st))
    eng.fwd_seq += 1
    torch._foreach_add_([bn.num_batches_tracked for bn in bns], 1)
    model._b200_epoch += 1  # running statistics changed behind autograd's back: eval-mode weights must be refolded
    return eng, logits


def train_backward(model, eng, dlogits, grads=None):
    """dlogits fp32, same shape as the logits of train_forward -> flat fp32 gradient (parameters() order)."""
    flat = flatten_parameters_(model)
    if grads is None:
        grads = torch.empty_like(flat)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.unet_b200_train_backward(eng.handle, dlogits.data_ptr(), flat.data_ptr(), grads.data_ptr(), st))
    return grads


class _UNetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x4, *params):
        eng, logits = train_forward(model, x4)
        ctx.model, ctx.eng, ctx.seq = model, eng, eng.fwd_seq
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        if ctx.eng.fwd_seq != ctx.seq:
            # the trainer keeps ONE set of saved activations per (batch, H, W): a second train-mode forward of the same shape
            # has overwritten the ones this backward needs
            raise RuntimeError("UNet (B200): backward() of a train-mode forward whose saved activations were overwritten by a "
                               "later forward of the same shape - call backward() before the next forward")
        grads = train_backward(model, ctx.eng, dlogits.contiguous().to(torch.float32))
        out, off = [], 0
        for p in model.parameters():
            n = p.numel()
            out.append(grads[off:off + n].view(p.shape))
            off += n
        return (None, None) + tuple(out)


def forward_train_autograd(model, x):
    """UNet.forward in training mode: float NCHW -> logits NCHW with a grad_fn that runs the B200 backward."""
    B, _, H, W = x.shape
    xin = x.detach().to(torch.float32).contiguous()
    x4 = torch.empty(B, H, W, 4, dtype=torch.bfloat16, device=x.device)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.unet_b200_nchw_to_nhwc4(xin.data_ptr(), B, model.in_channels, H, W, x4.data_ptr(), st))
    flatten_parameters_(model)
    logits = _UNetTrainFn.apply(model, x4, *model.parameters())
    return logits.reshape(B, model.out_channels, H, W).to(x.dtype)


def bce_dice_loss(logits, target, pos_weight=3.0, bce_weight=0.5, dice_weight=0.5, smooth=1e-6, want_grad=True):
    """Fused BCEDiceLoss (README.md:1868-1893). Returns (losses[3] = total,bce,dice on device, dlogits or None)."""
    z = logits.contiguous().to(torch.float32)
    t = target.contiguous().to(torch.float32)
    if z.numel() != t.numel():
        raise ValueError("logits and target must have the same number of elements")
    scratch = torch.empty(4, dtype=torch.float64, device=z.device)
    losses = torch.empty(3, dtype=torch.float32, device=z.device)
    dz = torch.empty_like(z) if want_grad else None
    check(lib.unet_b200_bce_dice_loss(z.data_ptr(), t.data_ptr(), z.numel(), float(pos_weight), float(bce_weight),
                                      float(dice_weight), float(smooth), scratch.data_ptr(), losses.data_ptr(),
                                      None if dz is None else dz.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return losses, dz


def allreduce_gradients(flat_grads, group=None):
    """Data-parallel exchange step: ONE all-reduce(sum) of the flat fp32 gradient buffer (NCCL over NVLink on GPUs, gloo in
    the CPU tests). Returns the factor the optimizer must apply to the summed gradient (1/world) - the scaling is folded
    into the AdamW kernel instead of a separate pass over the buffer."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


# ---------------------------------------------------------------------------------------------------- gradient buckets
MIN_BUCKET = 2 << 20      # elements: pending gradient ranges are exchanged once this much is final (8 MB of fp32) ...
MIN_INTERVAL = 1 << 16    # ... as one operation per contiguous range of at least this many elements; the rest waits


def plan_buckets(stage_ranges, min_bucket=MIN_BUCKET, min_interval=MIN_INTERVAL):
    """Cut the flat gradient into exchange buckets. stage_ranges[s] = flat range [lo, hi) that is final after backward stage
    s (unet_b200_trainer_stage_range). Returns [(stage, lo, hi)]: after `stage` the contiguous range [lo, hi) is exchanged.
    Ranges of consecutive stages that touch in memory are merged; nothing is sent before min_bucket elements are pending, and
    small leftovers (the 65-element head, final at stage 0 but stored last) ride with whichever neighbour completes next.
    Pure host arithmetic - identical on every rank."""
    pending, out = [], []
    last = len(stage_ranges) - 1
    for s, (lo, hi) in enumerate(stage_ranges):
        if hi > lo:
            pending.append([lo, hi])
        pending.sort()
        merged = []
        for iv in pending:
            if merged and merged[-1][1] == iv[0]:
                merged[-1][1] = iv[1]
            else:
                merged.append(list(iv))
        pending = merged
        if s == last or sum(b - a for a, b in pending) >= min_bucket:
            keep = []
            for a, b in pending:
                if s == last or b - a >= min_interval:
                    out.append((s, a, b))
                else:
                    keep.append([a, b])
            pending = keep
    return out


def shard_of(lo, hi, rank, world):
    """This rank's part [a, b) of bucket [lo, hi): `world` pieces whose starts stay multiples of 4 relative to lo."""
    piece = ((hi - lo + world - 1) // world + 3) // 4 * 4
    a = min(lo + rank * piece, hi)
    return a, min(a + piece, hi)


def _pick_exchange(group=None):
    """'nvlink' if every rank of `group` can map every other rank's memory (one host, same set of visible GPUs, one
    distinct device per rank), else 'nccl'. The same answer on every rank (decided from an all-gather)."""
    import os
    import socket
    if not torch.cuda.is_available() or dist.get_backend(group) != "nccl":
        return "nccl"   # i.e. dist.all_reduce on whatever backend the group has (gloo in the CPU tests)
    me = (socket.gethostname(), os.environ.get("CUDA_VISIBLE_DEVICES", ""), torch.cuda.current_device(), torch.cuda.device_count())
    world = dist.get_world_size(group)
    everyone = [None] * world
    dist.all_gather_object(everyone, me, group=group)
    same_view = len({(h, v, n) for h, v, _, n in everyone}) == 1
    distinct = len({d for _, _, d, _ in everyone}) == world
    return "nvlink" if same_view and distinct else "nccl"


class NvlinkExchange:
    """Gradient exchange without a collective call (one process per GPU, NVLink / NVSwitch peer access):

    * the flat gradient and the flat parameters live in torch symmetric memory, so every rank holds the peer-mapped base
      pointers (and, with NVLS, the multicast addresses) of all replicas;
    * pull / multimem: per gradient bucket, once every replica's bucket is final (device-side barrier on a side stream), ONE
      kernel per rank sums its part of the bucket over the replicas (peer loads, or multimem.ld_reduce inside the switch),
      applies AdamW on it (ZeRO-1: moments exist only for the owned parts) and stores the new parameters into every
      replica's buffer (peer stores / multimem.st) - reduce-scatter + optimizer + all-gather in one pass, overlapped with the
      rest of the backward;
    * push: every gradient atomic of the backward kernels goes straight to the owner rank (GradRoute in csrc/ptx.cuh), then a
      sharded AdamW; kept for comparison (slower: the 4-byte remote atomics are not coalesced), not bucketed."""

    def __init__(self, model, buckets, group=None, mode="pull"):
        import torch.distributed._symmetric_memory as symm_mem
        self.mode = mode
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        params = list(model.parameters())
        dev = params[0].device
        self.n = sum(p.numel() for p in params)
        self.shard = ((self.n + self.world - 1) // self.world + 3) // 4 * 4   # as unet_b200_adamw_step_p2p cuts it
        # parameters: symmetric, identical on every rank (rank 0's values win, as DDP does at construction)
        if getattr(model, "_b200_flat_symm", None) is None:
            model._b200_flat = None
            flat = flatten_parameters_(model, alloc=lambda n, d: symm_mem.empty(n, dtype=torch.float32, device=d))
            model._b200_flat_symm = symm_mem.rendezvous(flat, self.group)
        self.params = flatten_parameters_(model)
        self.hdl_p = model._b200_flat_symm
        dist.broadcast(self.params, src=dist.get_global_rank(self.group, 0), group=self.group)
        self.grads = symm_mem.empty(self.shard * self.world, dtype=torch.float32, device=dev)
        self.grads.zero_()
        self.hdl_g = symm_mem.rendezvous(self.grads, self.group)
        # NVLS: multicast mappings of both buffers, when the fabric offers them (0 otherwise -> peer loads / stores)
        off_g = self.grads.data_ptr() - int(self.hdl_g.buffer_ptrs[self.rank])
        off_p = self.params.data_ptr() - int(self.hdl_p.buffer_ptrs[self.rank])
        if off_g != 0 or off_p != 0:
            raise RuntimeError("symmetric buffers are expected to start at their allocation base")
        self.mc_g, self.mc_p = int(self.hdl_g.multicast_ptr), int(self.hdl_p.multicast_ptr)
        if mode == "auto":
            # measured on B200: peer loads / stores win at 2 GPUs (19.60 vs 19.93 ms), the switch wins at 8
            # (19.81 vs 20.1 ms); every rank must take the same branch, so the availability flag is min-reduced
            ok = torch.tensor([1 if (self.mc_g and self.mc_p and self.world >= 4) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            mode = "multimem" if int(ok.item()) else "pull"
            self.mode = mode
        if mode == "multimem" and (self.mc_g == 0 or self.mc_p == 0):
            raise RuntimeError("exchange='nvlink_mc' needs NVLS multicast support (symmetric memory multicast_ptr is 0)")
        # owned parts: (global lo, global hi, offset into the local moment arrays), one per bucket (push: one shard)
        if mode == "push":
            lo = min(self.rank * self.shard, self.n)
            self.parts = [(lo, min(lo + self.shard, self.n), 0)]
            self.buckets = []
        else:
            if any(lo % 4 for _, lo, _ in buckets):
                buckets = [(buckets[-1][0], 0, self.n)]     # unaligned tensor offsets: one bucket after the whole backward
            self.buckets = list(buckets)
            self.parts, off = [], 0
            for _, lo, hi in self.buckets:
                a, b = shard_of(lo, hi, self.rank, self.world)
                self.parts.append((a, b, off))
                off += (b - a + 3) // 4 * 4      # every part's moments start 16-byte aligned (the kernel moves float4)
        n_local = max(4, max((off_ + (b - a + 3) // 4 * 4) for a, b, off_ in self.parts))
        self.exp_avg = torch.zeros(n_local, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n_local, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    # ---- one bucket: barrier -> reduce-scatter + AdamW + all-gather kernel (runs on the CURRENT stream = the side stream)
    def exchange_bucket(self, idx, step_dev, lr, lr_dev, betas, eps, weight_decay):
        a, b, off = self.parts[idx]
        st = torch.cuda.current_stream().cuda_stream
        self.hdl_g.barrier(channel=0)      # every replica's gradient of this bucket is final
        if b <= a:
            return
        m, v = self.exp_avg[off:].data_ptr(), self.exp_avg_sq[off:].data_ptr()
        if self.mode == "multimem":
            check(lib.unet_b200_adamw_range_multimem(self.mc_p, self.mc_g, self.params.data_ptr(), a, b, m, v, float(lr),
                                                     lr_dev.data_ptr(), float(betas[0]), float(betas[1]), float(eps),
                                                     float(weight_decay), step_dev.data_ptr(), 1.0 / self.world, st))
        else:
            check(lib.unet_b200_adamw_range_p2p(self.hdl_p.buffer_ptrs_dev, self.hdl_g.buffer_ptrs_dev, self.world, self.rank,
                                                self.grads.data_ptr(), a, b, m, v, float(lr), lr_dev.data_ptr(), float(betas[0]),
                                                float(betas[1]), float(eps), float(weight_decay), step_dev.data_ptr(),
                                                1.0 / self.world, st))

    def close_step(self):
        # every replica holds the new parameters before the next forward reads them, and nobody still reads this replica's
        # gradient when the next backward clears it
        self.hdl_p.barrier(channel=1)

    # ---- push mode (whole backward, not bucketed)
    def backward_push(self, eng, dlogits):
        st = torch.cuda.current_stream().cuda_stream
        check(lib.unet_b200_train_backward_p2p(eng.handle, dlogits.data_ptr(), self.params.data_ptr(), self.grads.data_ptr(),
                                               self.hdl_g.buffer_ptrs_dev, self.world, st))
        self.hdl_g.barrier(channel=0)      # every replica's atomics have landed

    def optimizer_step_push(self, step_dev, lr, lr_dev, betas, eps, weight_decay):
        st = torch.cuda.current_stream().cuda_stream
        a, b, _ = self.parts[0]
        check(lib.unet_b200_adamw_range_p2p(self.hdl_p.buffer_ptrs_dev, None, self.world, self.rank, self.grads.data_ptr(), a, b,
                                            self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), float(lr), lr_dev.data_ptr(),
                                            float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                            step_dev.data_ptr(), 1.0 / self.world, st))
        self.hdl_p.barrier(channel=1)

    def reduced_parts(self):
        """[(lo, hi, sum over replicas of grads[lo:hi])] for the parts this rank owns, formed the way the exchange kernel forms
        it (peer loads / multimem.ld_reduce). Valid after a step until the next backward (the kernels do not modify the
        gradient buffers); push mode: only between backward_push and optimizer_step_push. Used by tests and bench.py."""
        out = []
        dev = self.grads.device
        for a, b, _ in self.parts:
            if b <= a:
                continue
            if self.mode == "push":
                out.append((a, b, self.grads[a:b].clone()))
            elif self.mode == "multimem":
                t = torch.empty(b - a, dtype=torch.float32, device=dev)
                check(lib.unet_b200_multimem_reduce(self.mc_g, a, b - a, t.data_ptr(), torch.cuda.current_stream().cuda_stream))
                out.append((a, b, t))
            else:
                acc = torch.zeros(b - a, dtype=torch.float32, device=dev)
                for r in range(self.world):
                    peer = self.hdl_g.get_buffer((self.rank + r) % self.world, (self.grads.numel(),), torch.float32)
                    acc += peer[a:b]
                out.append((a, b, acc))
        return out


class FusedTrainStep:
    """zero_grad -> forward -> BCEDiceLoss -> backward -> gradient exchange -> AdamW.step, README.md:2071-2079 with the
    criterion / optimizer of README.md:2169-2174, all on the B200 kernels. Data-parallel: pass a process group (or
    initialise torch.distributed); BatchNorm statistics and the loss stay per replica (the reference is single-device,
    SURVEY.md 8(e)).

    cuda_graph=True: the kernel launches of a step are captured once per (input shape, hyper-parameters) and replayed; the
    first step of a configuration runs eagerly, the second captures. Results are the same kernels either way. `lr` may be
    changed between steps (a scheduler): it lives in device memory, so the same graph serves every value."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, bce_weight=0.5, dice_weight=0.5,
                 pos_weight=3.0, smooth=1e-6, process_group=None, cuda_graph=True, exchange="auto", overlap=True,
                 bucket_elems=MIN_BUCKET):
        """exchange (world > 1):
        "nvlink"      = "nvlink_mc" when the fabric offers multicast and world >= 4, else "nvlink_pull";
        "nvlink_pull" no collective call: per gradient bucket ONE kernel sums the owned part over the peers' buffers (NVLink
                      loads), applies AdamW on it (ZeRO-1: moments exist only for the owned parts) and stores the new parameters
                      to all replicas (NVLink stores); device-side barriers, all of it inside the CUDA graph;
        "nvlink_mc"   the same kernel on NVSwitch multicast addresses: the switch sums the gradients (multimem.ld_reduce)
                      and broadcasts the new parameters (multimem.st); needs NVLS support;
        "nvlink_push" gradient atomics go to the owner replica inside the backward kernels instead (measured slower: the
                      4-byte remote atomics of the wgrad epilogues are not coalesced);
        "nccl"        per bucket an all-reduce of the flat gradient range + AdamW on the range, full optimizer state on every
                      replica;
        "auto"        nvlink when every rank of the group sits on its own visible GPU of one host (peer mapping possible),
                      else nccl (e.g. processes pinned with CUDA_VISIBLE_DEVICES to one device each).
        overlap: exchange every bucket on a side stream as soon as its backward stage has finished (default); False = one
        bucket after the whole backward (the round-1 behaviour, kept for A/B measurement). On ONE GPU there is nothing to
        exchange, but the AdamW update of a bucket still runs on the side stream beside the rest of the backward.
        bucket_elems: gradient elements that must be final before a bucket is sent (plan_buckets)."""
        if exchange not in ("auto", "nccl", "nvlink", "nvlink_pull", "nvlink_mc", "nvlink_push"):
            raise ValueError("exchange must be 'auto', 'nccl', 'nvlink', 'nvlink_pull', 'nvlink_mc' or 'nvlink_push'")
        self.exchange = exchange
        self.overlap = overlap
        self.bucket_elems = int(bucket_elems)
        self.nvlink = None
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.loss_cfg = dict(pos_weight=pos_weight, bce_weight=bce_weight, dice_weight=dice_weight, smooth=smooth)
        self.group = process_group
        self.cuda_graph = cuda_graph
        self.step_count = 0
        self.step_dev = None       # int32 device copy of step_count (read by the AdamW kernel)
        self.lr_dev = None         # fp32 device copy of lr (read by the AdamW kernel)
        self._lr_on_dev = None
        self.exp_avg = None
        self.exp_avg_sq = None
        self.grads = None          # persistent flat gradient buffer (world > 1)
        self.buckets = None
        self.last_grads = None
        self._side = None
        self._last_eng = None
        self._synced = False
        self._graphs = OrderedDict()
        self._seen = OrderedDict()

    MAX_GRAPHS = 2   # captured steps kept alive (each pins its private pool: logits, dz, static inputs)

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _sync_replicas(self, flat):
        """world > 1: every replica starts from rank 0's parameters and BatchNorm buffers (what DDP does at construction)."""
        if self._synced or self._world() == 1:
            return
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(flat, src=src, group=self.group)
        for b in self.model.buffers():
            dist.broadcast(b, src=src, group=self.group)
        self._synced = True

    def _prepare(self, images_shape, device):
        world = self._world()
        if self.exchange == "auto":
            self.exchange = "nccl" if world == 1 else _pick_exchange(self.group)
        staged = world > 1 or self.overlap     # one GPU: the optimizer of a bucket still runs beside the rest of the backward
        if staged and self.buckets is None and (world > 1 or images_shape is not None):
            B, H, W = images_shape
            eng = _engine(self.model, device, B, H, W)
            self.buckets = (plan_buckets(eng.stage_ranges, self.bucket_elems, min(MIN_INTERVAL, self.bucket_elems))
                            if self.overlap else [(eng.n_stages - 1, 0, eng.n_params)])
        if self.step_dev is None:
            self.step_dev = torch.full((1,), self.step_count, dtype=torch.int32, device=device)
            self.lr_dev = torch.full((1,), float(self.lr), dtype=torch.float32, device=device)
            self._lr_on_dev = float(self.lr)
        if self.exchange != "nccl" and self.nvlink is None and world > 1:
            mode = {"nvlink_push": "push", "nvlink_mc": "multimem", "nvlink_pull": "pull"}.get(self.exchange, "auto")
            self.nvlink = NvlinkExchange(self.model, self.buckets, self.group, mode=mode)
            self.exchange = {"multimem": "nvlink_mc", "pull": "nvlink", "push": "nvlink_push"}[self.nvlink.mode]
            self.buckets = self.nvlink.buckets
            self.exp_avg, self.exp_avg_sq = self.nvlink.exp_avg, self.nvlink.exp_avg_sq
            self.grads = self.nvlink.grads
            self._synced = True        # NvlinkExchange broadcast the parameters
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            for b in self.model.buffers():
                dist.broadcast(b, src=src, group=self.group)
        flat = flatten_parameters_(self.model)
        if self.nvlink is None:
            self._sync_replicas(flat)
            if self.exp_avg is None or self.exp_avg.numel() != flat.numel() or self.exp_avg.device != flat.device:
                self.exp_avg = torch.zeros_like(flat)
                self.exp_avg_sq = torch.zeros_like(flat)
            if staged and (self.grads is None or self.grads.numel() != flat.numel()):
                self.grads = torch.empty_like(flat)
        if staged and self._side is None:
            self._side = torch.cuda.Stream(device=device)
        if self._lr_on_dev != float(self.lr):
            self.lr_dev.fill_(float(self.lr))
            self._lr_on_dev = float(self.lr)
        return flat

    def _adamw_range(self, flat, grads, lo, hi, grad_scale):
        check(lib.unet_b200_adamw_step_dev(flat[lo:].data_ptr(), grads[lo:].data_ptr(), self.exp_avg[lo:].data_ptr(),
                                           self.exp_avg_sq[lo:].data_ptr(), hi - lo, float(self.lr), self.lr_dev.data_ptr(),
                                           float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                           self.step_dev.data_ptr(), grad_scale, torch.cuda.current_stream().cuda_stream))

    def _run(self, images, masks):
        """All kernels of one step on the current stream (eager or under graph capture)."""
        model = self.model
        if images.dtype == torch.bfloat16 and images.dim() == 4 and images.shape[-1] == 4:
            x4 = images
        else:
            B, _, H, W = images.shape
            xin = images if images.dtype == torch.float32 else images.to(torch.float32)
            x4 = torch.empty(B, H, W, 4, dtype=torch.bfloat16, device=images.device)
            check(lib.unet_b200_nchw_to_nhwc4(xin.data_ptr(), B, model.in_channels, H, W, x4.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        flat = flatten_parameters_(model)
        eng, logits = train_forward(model, x4)
        self._last_eng = eng
        losses, dz = bce_dice_loss(logits, masks.reshape(logits.shape), **self.loss_cfg)
        self.step_dev.add_(1)
        world = self._world()
        if world == 1 and not self.overlap:
            grads = train_backward(model, eng, dz)
            self._adamw_range(flat, grads, 0, flat.numel(), 1.0)
            return losses, grads
        if self.nvlink is not None and self.nvlink.mode == "push":
            self.nvlink.backward_push(eng, dz)
            self.nvlink.optimizer_step_push(self.step_dev, self.lr, self.lr_dev, self.betas, self.eps, self.weight_decay)
            return losses, self.grads
        # staged backward on the main stream; each bucket's exchange + AdamW on the side stream once its stages are done
        main = torch.cuda.current_stream()
        side = self._side
        grads = self.grads
        nb = 0
        for s in range(eng.n_stages):
            check(lib.unet_b200_train_backward_stage(eng.handle, s, dz.data_ptr(), flat.data_ptr(), grads.data_ptr(), main.cuda_stream))
            while nb < len(self.buckets) and self.buckets[nb][0] == s:
                _, lo, hi = self.buckets[nb]
                check(lib.unet_b200_trainer_join(eng.handle, main.cuda_stream, side.cuda_stream))
                with torch.cuda.stream(side):
                    if self.nvlink is not None:
                        self.nvlink.exchange_bucket(nb, self.step_dev, self.lr, self.lr_dev, self.betas, self.eps, self.weight_decay)
                    else:
                        if world > 1:
                            dist.all_reduce(grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                        self._adamw_range(flat, grads, lo, hi, 1.0 / world)
                nb += 1
        with torch.cuda.stream(side):
            if self.nvlink is not None:
                self.nvlink.close_step()
        main.wait_stream(side)
        return losses, grads

    def step(self, images, masks):
        """images: float NCHW [B,3,H,W] (or bf16 NHWC4 [B,H,W,4]); masks: [B,1,H,W] / [B,H,W] in {0,1}.
        Returns a device tensor [total, bce, dice] (no host sync; with cuda_graph it is overwritten by the next step)."""
        model = self.model
        if not model.training:
            raise RuntimeError("FusedTrainStep.step needs model.train()")
        if not images.is_cuda or not masks.is_cuda:
            raise RuntimeError("FusedTrainStep (B200): inputs must be CUDA tensors - there is no CPU fallback")
        images = images.detach().contiguous()
        masks = masks.detach().contiguous().to(torch.float32)
        if images.dim() == 4 and images.dtype == torch.bfloat16 and images.shape[-1] == 4:
            bhw = (images.shape[0], images.shape[1], images.shape[2])
        else:
            bhw = (images.shape[0], images.shape[2], images.shape[3])
        flat = self._prepare(bhw, images.device)
        # (lr is not part of the key: it is read from device memory)
        key = (tuple(images.shape), images.dtype, tuple(masks.shape), flat.data_ptr(), float(self.weight_decay),
               tuple(self.betas), float(self.eps))
        self.step_count += 1
        use_graph = self.cuda_graph
        if not use_graph or key not in self._seen:
            self._seen[key] = True                 # first step of a configuration: eager (creates engines, sets attributes)
            while len(self._seen) > 4 * self.MAX_GRAPHS:
                self._seen.popitem(last=False)
            losses, self.last_grads = self._run(images, masks)
        else:
            entry = self._graphs.get(key)
            if entry is None:
                while len(self._graphs) >= self.MAX_GRAPHS:
                    self._graphs.popitem(last=False)
                sx, sy = torch.empty_like(images), torch.empty_like(masks)
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=images.device)):   # capture stream on THIS device
                    out = self._run(sx, sy)
                # the graph replays into the trainer's workspace: the entry keeps the engine alive even if the model's engine
                # cache lets go of it
                entry = (graph, sx, sy, out, self._last_eng)
                self._graphs[key] = entry
            else:
                self._graphs.move_to_end(key)
            graph, sx, sy, (losses, grads), _ = entry
            if sx.data_ptr() != images.data_ptr():
                sx.copy_(images, non_blocking=True)
            if sy.data_ptr() != masks.data_ptr():
                sy.copy_(masks, non_blocking=True)
            graph.replay()
            self.last_grads = grads
        model._b200_epoch += 1  # kernels wrote parameters / BN buffers behind autograd's back: repack before the next eval
        return losses

    # ------------------------------------------------------------------ optimizer state in torch.optim.AdamW's format
    def _full_moments(self):
        """(exp_avg, exp_avg_sq) as full flat tensors. NVLink exchange: the moments are sharded over the replicas (ZeRO-1), so
        this is a COLLECTIVE - every rank must call it."""
        if self.nvlink is None:
            return self.exp_avg, self.exp_avg_sq
        nv = self.nvlink
        full = torch.zeros(2, nv.n, dtype=torch.float32, device=self.exp_avg.device)
        for a, b, off in nv.parts:
            full[0, a:b] = self.exp_avg[off:off + b - a]
            full[1, a:b] = self.exp_avg_sq[off:off + b - a]
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)   # the parts are disjoint: the sum is the gather
        return full[0], full[1]

    def state_dict(self):
        """Same structure torch.optim.AdamW.state_dict() produces for optimizer = AdamW(model.parameters(), ...)
        (the 'optimizer_state_dict' of README.md:2208-2213), so either optimizer can resume from the other's file.
        With an NVLink exchange this gathers the sharded moments: call it on EVERY rank (write the file on one)."""
        params = list(self.model.parameters())
        state = {}
        if self.exp_avg is not None and self.step_count > 0:
            exp_avg, exp_avg_sq = self._full_moments()
            off = 0
            for i, p in enumerate(params):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": exp_avg[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": exp_avg_sq[off:off + n].view(p.shape).clone()}
                off += n
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": True, "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd, images_shape=None):
        """Restore lr / betas / eps / weight_decay, the step count and the moments. With an NVLink exchange every rank keeps
        only the ranges it owns. images_shape = (B, H, W) of the training batches is needed when the exchange has not been
        set up yet (world > 1 and no step taken): the bucket plan comes from the trainer of that shape."""
        group = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = group["lr"], tuple(group["betas"]), group["eps"], group["weight_decay"]
        params = list(self.model.parameters())
        if self._world() > 1 and self.buckets is None and images_shape is None:
            raise ValueError("load_state_dict before the first step of a data-parallel run needs images_shape=(B, H, W)")
        flat = self._prepare(images_shape, params[0].device)
        dev = flat.device
        full = torch.zeros(2, flat.numel(), dtype=torch.float32, device=dev)
        self.step_count = 0
        off = 0
        for i, p in enumerate(params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                full[0, off:off + n] = st["exp_avg"].reshape(-1).to(dev)
                full[1, off:off + n] = st["exp_avg_sq"].reshape(-1).to(dev)
                self.step_count = int(st["step"])
            off += n
        if self.nvlink is None:
            self.exp_avg.copy_(full[0])
            self.exp_avg_sq.copy_(full[1])
        else:
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for a, b, loc in self.nvlink.parts:      # the global range [a, b) lives at [loc, loc + b - a) of the local arrays
                self.exp_avg[loc:loc + b - a] = full[0, a:b]
                self.exp_avg_sq[loc:loc + b - a] = full[1, a:b]
        self.step_dev.fill_(self.step_count)
        self._graphs.clear()
