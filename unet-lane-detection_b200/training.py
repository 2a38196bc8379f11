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


def flatten_parameters_(model):
    """Make every parameter of `model` a view into one contiguous fp32 buffer (parameters() order) and return it.
    Parameter objects keep their identity, so optimizers created before or after stay valid."""
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
    flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
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


class FusedTrainStep:
    """zero_grad -> forward -> BCEDiceLoss -> backward -> (all-reduce) -> AdamW.step, README.md:2071-2079 with the
    criterion / optimizer of README.md:2169-2174, all on the B200 kernels. Data-parallel: pass a process group (or
    initialise torch.distributed) and gradients are summed over ranks with ONE NCCL all-reduce of the flat buffer;
    BatchNorm statistics and the loss stay per replica (the reference is single-device).

    cuda_graph=True: the ~270 kernel launches of a step are captured once per (input shape, hyper-parameters) and
    replayed; the first step of a configuration runs eagerly, the second captures. Results are the same kernels
    either way. `lr` may be changed between steps (a scheduler): the graph is re-captured for the new value."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, bce_weight=0.5, dice_weight=0.5,
                 pos_weight=3.0, smooth=1e-6, process_group=None, cuda_graph=True):
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
        self._graphs = {}
        self._seen = set()

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _prepare(self, device):
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
        if not self.cuda_graph or key not in self._seen:
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
        if self.exp_avg is not None and self.step_count > 0:
            off = 0
            for i, p in enumerate(params):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
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
