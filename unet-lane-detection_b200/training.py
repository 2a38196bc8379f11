"""Training step on the B200 kernels (reference: README.md:2060-2084 train_one_epoch, 1855-1893 BCEDiceLoss,
2173-2174 AdamW).

Two ways in, both through libunet_b200.so (no PyTorch/cuDNN compute, no CPU fallback):

* drop-in autograd: `model.train(); out = model(images); loss = criterion(out, masks); loss.backward();
  optimizer.step()` works unchanged - `UNet.forward` in training mode runs `unet_b200_train_forward` inside a
  `torch.autograd.Function` whose backward is `unet_b200_train_backward`, so the reference's own BCEDiceLoss and
  `torch.optim.AdamW` can stay.
* fused step: `FusedTrainStep(model).step(images, masks)` additionally runs the loss (+ its gradient), the
  data-parallel gradient all-reduce (NCCL, one flat buffer) and AdamW as hand-written kernels.

Parameters are kept in ONE flat fp32 buffer in `model.parameters()` order; every nn.Parameter is a view into it, so
state_dict / optimizers see ordinary tensors.
"""
import ctypes as C

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


def _engine(model, device, B, H, W):
    key = ("train", str(device), B, H, W)
    eng = model._engines.get(key)
    if eng is None:
        eng = _TrainEngine(model, device, B, H, W)
        model._engines[key] = eng
        sizes = [p.numel() for p in model.parameters()]
        want = [eng.offsets[i + 1] - eng.offsets[i] for i in range(len(eng.offsets) - 1)]
        if sizes != want:
            raise RuntimeError("UNet (B200): parameter layout of the module does not match the library's trainer")
    return eng


def _bn_tables(model):
    bns = [bn for _, bn in model._double_convs()]
    n = len(bns)
    means = (C.c_void_p * n)(*[bn.running_mean.data_ptr() for bn in bns])
    vars_ = (C.c_void_p * n)(*[bn.running_var.data_ptr() for bn in bns])
    return bns, means, vars_


def train_forward(model, x4):
    """x4: bf16 NHWC4 [B,H,W,4] -> logits fp32 [B,H,W]; updates BN running statistics like nn.BatchNorm2d."""
    B, H, W, _ = x4.shape
    eng = _engine(model, x4.device, B, H, W)
    flat = flatten_parameters_(model)
    bns, means, vars_ = _bn_tables(model)
    logits = torch.empty(B, H, W, dtype=torch.float32, device=x4.device)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.unet_b200_train_forward(eng.handle, x4.data_ptr(), flat.data_ptr(), means, vars_, float(bns[0].momentum),
                                      float(bns[0].eps), logits.data_ptr(), st))
    torch._foreach_add_([bn.num_batches_tracked for bn in bns], 1)
    model._b200_epoch += 1  # running statistics changed behind autograd's back: eval-mode weights must be refolded
    return eng, logits


def train_backward(model, eng, dlogits):
    """dlogits fp32 [B,H,W] -> flat fp32 gradient (parameters() order)."""
    flat = flatten_parameters_(model)
    grads = torch.empty_like(flat)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.unet_b200_train_backward(eng.handle, dlogits.data_ptr(), flat.data_ptr(), grads.data_ptr(), st))
    return grads


class _UNetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x4, *params):
        eng, logits = train_forward(model, x4)
        ctx.model, ctx.eng = model, eng
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
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
    return logits.reshape(B, 1, H, W).to(x.dtype)


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
      pointers of all replicas;
    * backward: every gradient atomic goes straight to the OWNER rank's gradient shard (GradRoute in csrc/ptx.cuh) -
      the reduce-scatter is fused into the wgrad / BN / bias-gradient kernels;
    * device-side barrier (symmetric-memory signal pads, on the stream);
    * AdamW on the owned shard only (ZeRO-1: moments exist only for the shard) whose stores go to ALL replicas'
      parameter buffers - the all-gather is fused into the optimizer kernel; then a second barrier.

    Per step and GPU: (world-1)/world of 124 MB leaves over NVLink during the backward and the same again during the
    optimizer, fully overlapped with compute; NCCL's ring all-reduce moves twice that after the backward has finished."""

    def __init__(self, model, group=None, mode="pull"):
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
        self.exp_avg = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def backward(self, model, eng, dlogits):
        st = torch.cuda.current_stream().cuda_stream
        if self.mode == "push":
            check(lib.unet_b200_train_backward_p2p(eng.handle, dlogits.data_ptr(), self.params.data_ptr(), self.grads.data_ptr(),
                                                   self.hdl_g.buffer_ptrs_dev, self.world, st))
        else:
            check(lib.unet_b200_train_backward(eng.handle, dlogits.data_ptr(), self.params.data_ptr(), self.grads.data_ptr(), st))
        self.hdl_g.barrier(channel=0)      # push: every replica's atomics have landed; pull: every replica's gradient is complete

    def optimizer_step(self, step_dev, lr, betas, eps, weight_decay):
        st = torch.cuda.current_stream().cuda_stream
        if self.mode == "multimem":
            check(lib.unet_b200_adamw_step_multimem(self.mc_p, self.mc_g, self.params.data_ptr(), self.world, self.rank,
                                                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n, float(lr),
                                                    float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                                    step_dev.data_ptr(), 1.0 / self.world, st))
            self.hdl_p.barrier(channel=1)
            return
        check(lib.unet_b200_adamw_step_p2p(self.hdl_p.buffer_ptrs_dev, self.hdl_g.buffer_ptrs_dev if self.mode == "pull" else None,
                                           self.world, self.rank, self.grads.data_ptr(),
                                           self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n, float(lr), float(betas[0]),
                                           float(betas[1]), float(eps), float(weight_decay), step_dev.data_ptr(),
                                           1.0 / self.world, st))
        # every replica holds the new parameters before the next forward reads them (pull: and nobody still reads this
        # replica's gradient when the next backward clears it)
        self.hdl_p.barrier(channel=1)

    def reduced_shard(self):
        """Sum over replicas of this rank's gradient shard, valid between backward() and optimizer_step() (tests)."""
        lo, hi = self.rank * self.shard, (self.rank + 1) * self.shard
        if self.mode == "push":
            return self.grads[lo:hi].clone()
        if self.mode == "multimem":    # the sum the optimizer kernel will see: formed by the switch
            out = torch.empty(self.shard, dtype=torch.float32, device=self.grads.device)
            check(lib.unet_b200_multimem_reduce(self.mc_g, lo, self.shard, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            return out
        full = [torch.empty_like(self.grads) for _ in range(self.world)]
        dist.all_gather(full, self.grads, group=self.group)
        return torch.stack(full).sum(0)[lo:hi]


class FusedTrainStep:
    """zero_grad -> forward -> BCEDiceLoss -> backward -> (all-reduce) -> AdamW.step, README.md:2071-2079 with the
    criterion / optimizer of README.md:2169-2174, all on the B200 kernels. Data-parallel: pass a process group (or
    initialise torch.distributed) and gradients are summed over ranks with ONE NCCL all-reduce of the flat buffer;
    BatchNorm statistics and the loss stay per replica (the reference is single-device).

    cuda_graph=True: the ~270 kernel launches of a step are captured once per (input shape, hyper-parameters) and
    replayed; the first step of a configuration runs eagerly, the second captures. Results are the same kernels
    either way. `lr` may be changed between steps (a scheduler): the graph is re-captured for the new value."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, bce_weight=0.5, dice_weight=0.5,
                 pos_weight=3.0, smooth=1e-6, process_group=None, cuda_graph=True, exchange="auto"):
        """exchange (world > 1):
        "nvlink"      = "nvlink_mc" when the fabric offers multicast and world >= 4, else "nvlink_pull";
        "nvlink_pull" no collective call: ONE kernel sums the owned gradient shard over the peers' buffers (NVLink loads),
                      applies AdamW on it (ZeRO-1: moments exist only for the shard) and stores the new parameters to all
                      replicas (NVLink stores); two device-side barriers per step, all of it inside the CUDA graph;
        "nvlink_mc"   the same kernel on NVSwitch multicast addresses: the switch sums the gradient shard (multimem.ld_reduce)
                      and broadcasts the new parameters (multimem.st); needs NVLS support;
        "nvlink_push" gradient atomics go to the owner replica inside the backward kernels instead (measured slower: the
                      4-byte remote atomics of the wgrad epilogues are not coalesced);
        "nccl"        one all-reduce of the flat gradient after the backward, full AdamW on every replica;
        "auto"        nvlink when every rank of the group sits on its own visible GPU of one host (peer mapping possible),
                      else nccl (e.g. processes pinned with CUDA_VISIBLE_DEVICES to one device each)."""
        if exchange not in ("auto", "nccl", "nvlink", "nvlink_pull", "nvlink_mc", "nvlink_push"):
            raise ValueError("exchange must be 'auto', 'nccl', 'nvlink', 'nvlink_pull', 'nvlink_mc' or 'nvlink_push'")
        self.exchange = exchange
        self.nvlink = None
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.loss_cfg = dict(pos_weight=pos_weight, bce_weight=bce_weight, dice_weight=dice_weight, smooth=smooth)
        self.group = process_group
        self.cuda_graph = cuda_graph
        self.step_count = 0
        self.step_dev = None       # int32 device copy of step_count (read by the AdamW kernel)
        self.exp_avg = None
        self.exp_avg_sq = None
        self.last_grads = None
        self.keep_grad_shard = False   # nvlink exchange: keep a copy of the reduced gradient shard of every step (tests)
        self.last_grad_shard = None
        self._graphs = {}
        self._seen = set()

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _prepare(self, device):
        if self.exchange == "auto":
            self.exchange = "nccl" if self._world() == 1 else _pick_exchange(self.group)
        if self.exchange != "nccl" and self.nvlink is None and self._world() > 1:
            self.nvlink = NvlinkExchange(self.model, self.group, mode={"nvlink_push": "push", "nvlink_mc": "multimem", "nvlink_pull": "pull"}.get(self.exchange, "auto"))
            self.exchange = {"multimem": "nvlink_mc", "pull": "nvlink", "push": "nvlink_push"}[self.nvlink.mode]
            self.step_dev = torch.full((1,), self.step_count, dtype=torch.int32, device=device)
            self.exp_avg, self.exp_avg_sq = self.nvlink.exp_avg, self.nvlink.exp_avg_sq
        if self.nvlink is not None:
            return flatten_parameters_(self.model)
        flat = flatten_parameters_(self.model)
        if self.exp_avg is None or self.exp_avg.numel() != flat.numel() or self.exp_avg.device != flat.device:
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
            self.step_dev = torch.full((1,), self.step_count, dtype=torch.int32, device=device)
        return flat

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
        losses, dz = bce_dice_loss(logits, masks.reshape(logits.shape), **self.loss_cfg)
        if self.nvlink is not None:
            self.nvlink.backward(model, eng, dz)
            if self.keep_grad_shard:
                self.last_grad_shard = self.nvlink.reduced_shard()
            self.step_dev.add_(1)
            self.nvlink.optimizer_step(self.step_dev, self.lr, self.betas, self.eps, self.weight_decay)
            return losses, self.nvlink.grads
        grads = train_backward(model, eng, dz)
        grad_scale = allreduce_gradients(grads, self.group)
        self.step_dev.add_(1)
        check(lib.unet_b200_adamw_step_dev(flat.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                           flat.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                           float(self.weight_decay), self.step_dev.data_ptr(), grad_scale,
                                           torch.cuda.current_stream().cuda_stream))
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
        flat = self._prepare(images.device)
        key = (tuple(images.shape), images.dtype, tuple(masks.shape), flat.data_ptr(), float(self.lr), float(self.weight_decay),
               tuple(self.betas), float(self.eps))
        self.step_count += 1
        use_graph = self.cuda_graph
        if not use_graph or key not in self._seen:
            self._seen.add(key)                    # first step of a configuration: eager (creates engines, sets attributes)
            losses, self.last_grads = self._run(images, masks)
        else:
            entry = self._graphs.get(key)
            if entry is None:
                sx, sy = torch.empty_like(images), torch.empty_like(masks)
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(graph):
                    out = self._run(sx, sy)
                entry = (graph, sx, sy, out)
                self._graphs[key] = entry
            graph, sx, sy, (losses, grads) = entry
            if sx.data_ptr() != images.data_ptr():
                sx.copy_(images, non_blocking=True)
            if sy.data_ptr() != masks.data_ptr():
                sy.copy_(masks, non_blocking=True)
            graph.replay()
            self.last_grads = grads
        model._b200_epoch += 1  # kernels wrote parameters / BN buffers behind autograd's back: repack before the next eval
        return losses

    # ------------------------------------------------------------------ optimizer state in torch.optim.AdamW's format
    def state_dict(self):
        """Same structure torch.optim.AdamW.state_dict() produces for optimizer = AdamW(model.parameters(), ...)
        (the 'optimizer_state_dict' of README.md:2208-2213), so either optimizer can resume from the other's file."""
        params = list(self.model.parameters())
        state = {}
        exp_avg, exp_avg_sq = self.exp_avg, self.exp_avg_sq
        if self.nvlink is not None:   # moments are sharded over the replicas: gather them for the checkpoint
            full = [torch.empty(self.nvlink.shard * self.nvlink.world, dtype=torch.float32, device=exp_avg.device) for _ in range(2)]
            dist.all_gather_into_tensor(full[0], self.exp_avg, group=self.group)
            dist.all_gather_into_tensor(full[1], self.exp_avg_sq, group=self.group)
            exp_avg, exp_avg_sq = full
        if exp_avg is not None and self.step_count > 0:
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

    def load_state_dict(self, sd):
        group = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = group["lr"], tuple(group["betas"]), group["eps"], group["weight_decay"]
        params = list(self.model.parameters())
        flat = self._prepare(params[0].device)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0
        off = 0
        for i, p in enumerate(params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1).to(flat.device))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(flat.device))
                self.step_count = int(st["step"])
            off += n
        self.step_dev.fill_(self.step_count)
        self._graphs.clear()
