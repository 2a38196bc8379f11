"""Drop-in `UNet` (reference README.md:1421-1481) whose eval forward runs on libunet_b200.so.

Same constructor signature, same submodule / parameter / buffer names, shapes, dtypes and
registration order (so state_dict()/load_state_dict() interchange with reference checkpoints,
SURVEY.md Appendix A), same forward contract: float NCHW [B,in_channels,H,W] -> logits
[B,out_channels,H,W]. The parameters live in ordinary nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d
containers, but those modules' own forward() is never called: forward() folds BN into the weights,
packs them for the tensor-core kernels and runs the hand-written CUDA path. CPU tensors, or a
missing/unsupported device, raise - there is no PyTorch/cuDNN fallback.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from ._lib import check, f3, lib
from .ops import MEAN_255, STD_255

PRECISIONS = {"bf16": 0, "fp32": 1}   # UB_PRECISION_* of include/unet_b200.h
MAX_ENGINES = 4   # bound plans kept per model (each owns a workspace + packed weights): least recently used goes first


class _Engine:
    """One bound plan: (device, batch capacity, H, W, precision) + workspace + packed weights."""

    def __init__(self, model: "UNet", device: torch.device, cap: int, H: int, W: int, precision: str = "bf16"):
        feats = (C.c_int * len(model.features))(*model.features)
        handle = C.c_void_p()
        check(lib.unet_b200_plan_create_ex(C.byref(handle), cap, H, W, model.in_channels, model.out_channels, feats,
                                           len(model.features), PRECISIONS[precision]))
        self.handle = handle
        self.cap, self.H, self.W, self.device, self.precision = cap, H, W, device, precision
        self.workspace = torch.empty(lib.unet_b200_plan_workspace_bytes(handle), dtype=torch.uint8, device=device)
        self.weights = torch.zeros(lib.unet_b200_plan_weight_bytes(handle), dtype=torch.uint8, device=device)
        check(lib.unet_b200_plan_bind(handle, self.workspace.data_ptr(), self.weights.data_ptr()))
        self.weights_key = None
        self.launches = lib.unet_b200_forward_launches(handle)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            lib.unet_b200_plan_destroy(h)
            self.handle = None


class UNet(nn.Module):
    """U-Net for lane segmentation; signature of the reference's class (README.md:1424)."""

    def __init__(self, in_channels=3, out_channels=1, features=[64, 128, 256, 512]):  # noqa: B006 (reference signature)
        super().__init__()
        self.in_channels, self.out_channels, self.features = in_channels, out_channels, list(features)
        self.encoder_blocks = nn.ModuleList()
        self.decoder_blocks = nn.ModuleList()
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        c = in_channels
        for f in self.features:
            self.encoder_blocks.append(self._conv_block(c, f))
            c = f
        self.bottleneck = self._conv_block(self.features[-1], self.features[-1] * 2)
        for f in reversed(self.features):
            self.decoder_blocks.append(nn.ConvTranspose2d(f * 2, f, kernel_size=2, stride=2))
            self.decoder_blocks.append(self._conv_block(f * 2, f))
        self.output = nn.Conv2d(self.features[0], out_channels, kernel_size=1)
        # B200 runtime state (not part of the state_dict)
        self.b200_chunk = 128  # frames per pass through the plan (bounds the workspace; tune for L2 reuse)
        # eval precision of forward / predict_mask / infer_host: "bf16" (fast path, logits within 2e-2 of the fp32 reference) or
        # "fp32" (a UB_PRECISION_FP32 plan: split-bf16 x3 on the same tensor-core kernel, logits within 1e-4; about 3x the
        # tensor work) - BASELINE.json north_star parity gates
        self.b200_precision = "bf16"
        self._engines = OrderedDict()   # LRU, at most MAX_ENGINES entries (shared with the trainers of training.py)
        self._b200_epoch = 0  # bumped whenever a kernel updates parameters / BN buffers in place
        # True: the weights will not change any more (a deployed model, e.g. inside B200_model_container) - the per-call
        # check that walks all 118 tensors for in-place updates is skipped once a plan has packed them
        self.b200_frozen = False
        self._last_engine = None
        self.gpu_launches = 0

    # copy.deepcopy(model) / torch.save(model): the plans, their workspaces and the flat training buffers belong to THIS object
    # (ctypes handles, CUDA graphs) - a copy starts without them and builds its own on first use
    _RUNTIME_STATE = ("_engines", "_last_engine", "_b200_key_slots", "_b200_flat", "_b200_flat_symm")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._RUNTIME_STATE:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._engines = OrderedDict()
        self._last_engine = None

    def _conv_block(self, in_channels, out_channels):
        return nn.Sequential(
            nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, 3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )

    # ------------------------------------------------------------------ engine / weights
    def _double_convs(self):
        """3x3 convs with their BatchNorms in plan order (include/unet_b200.h: plan_num_convs)."""
        blocks = list(self.encoder_blocks) + [self.bottleneck] + [self.decoder_blocks[i] for i in range(1, len(self.decoder_blocks), 2)]
        for blk in blocks:
            yield blk[0], blk[1]
            yield blk[3], blk[4]

    def _weights_key(self):
        """Identity + in-place version of every parameter and buffer: a changed key means the plan's packed weights are stale.
        Runs on every eval call of a live (not frozen) module, so the module tree is walked once and the key is read through
        the registered (dict, name) slots - a replaced Parameter object is still seen, the generators of `parameters()` /
        `buffers()` (most of the former 0.14 ms) are not re-run."""
        slots = self.__dict__.get("_b200_key_slots")
        if slots is None:
            slots = [(m._parameters, n) for m in self.modules() for n, p in m._parameters.items() if p is not None]
            slots += [(m._buffers, n) for m in self.modules() for n, b in m._buffers.items() if b is not None]
            self.__dict__["_b200_key_slots"] = slots
        return (self._b200_epoch,) + tuple((d[n].data_ptr(), d[n]._version) for d, n in slots)

    def _engine(self, device, H, W, batch, pipelined=False):
        prec = self.b200_precision
        if prec not in PRECISIONS:
            raise ValueError("b200_precision must be 'bf16' or 'fp32'")
        cap = min(max(1, int(self.b200_chunk)), batch) if batch < self.b200_chunk else int(self.b200_chunk)
        if pipelined == "halved" and batch >= 64 and cap >= batch:
            # host-buffer entry of a plan that cannot run in pieces: a batch that fits one pass is still cut in two, so that
            # the H2D copy of the second half and the D2H copy of the first overlap the kernels
            cap = (batch + 1) // 2
        key = (str(device), cap, H, W, prec)
        eng = self._engines.get(key)
        if eng is None:
            while len(self._engines) >= MAX_ENGINES:
                self._engines.popitem(last=False)
            eng = _Engine(self, device, cap, H, W, prec)
            self._engines[key] = eng
        else:
            self._engines.move_to_end(key)
        if self.b200_frozen and eng.weights_key is not None:
            return eng
        wkey = self._weights_key()
        if eng.weights_key != wkey:
            self._pack(eng)
            eng.weights_key = wkey
        return eng

    def _pack(self, eng):
        st = torch.cuda.current_stream().cuda_stream
        keep = []

        def dev32(t):
            t = t.detach().to(device=eng.device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        for i, (conv, bn) in enumerate(self._double_convs()):
            check(lib.unet_b200_plan_set_conv(eng.handle, i, dev32(conv.weight), dev32(bn.weight), dev32(bn.bias),
                                              dev32(bn.running_mean), dev32(bn.running_var), float(bn.eps), st))
        for j in range(len(self.features)):
            up = self.decoder_blocks[2 * j]
            check(lib.unet_b200_plan_set_convT(eng.handle, j, dev32(up.weight), dev32(up.bias), st))
        check(lib.unet_b200_plan_set_head(eng.handle, dev32(self.output.weight.reshape(self.out_channels, -1)),
                                          dev32(self.output.bias), st))
        torch.cuda.current_stream().synchronize()

    def _check_input(self, t, what):
        if not t.is_cuda:
            raise RuntimeError(f"UNet (B200): {what} is on {t.device}; the B200 path runs on CUDA sm_100 only and has "
                               "no CPU fallback - move the model and input with .to('cuda')")

    # ------------------------------------------------------------------ forward paths
    def forward(self, x):
        """Forward on the B200 kernels: float NCHW -> logits NCHW (README.md:1460-1481).
        eval(): BatchNorm folded into the weights. train(): batch-statistics BatchNorm, activations kept, and the
        result carries a grad_fn whose backward runs the B200 backward kernels (training.py)."""
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        if torch.jit.is_tracing():
            # torch.jit.trace cannot see a C-ABI call: the traced graph holds ONE operator, unet_b200::infer (torchscript.py)
            if self.training:
                raise RuntimeError("UNet (B200): trace the model in eval() mode")
            from .torchscript import traced_forward
            return traced_forward(self, x)
        self._check_input(x, "input")
        if self.training:
            from .training import forward_train_autograd
            return forward_train_autograd(self, x)
        B, _, H, W = x.shape
        xin = x.detach().to(torch.float32).contiguous()
        st = torch.cuda.current_stream().cuda_stream
        if self.b200_precision == "fp32":     # fp32-class plan: the image stays fp32
            x4 = torch.empty(B, H, W, 4, dtype=torch.float32, device=x.device)
            check(lib.unet_b200_nchw_to_nhwc4_f32(xin.data_ptr(), B, self.in_channels, H, W, x4.data_ptr(), st))
        else:
            x4 = torch.empty(B, H, W, 4, dtype=torch.bfloat16, device=x.device)
            check(lib.unet_b200_nchw_to_nhwc4(xin.data_ptr(), B, self.in_channels, H, W, x4.data_ptr(), st))
        self.gpu_launches += 1
        logits = self.forward_nhwc4(x4, want=("logits",))[0]
        return logits.reshape(B, self.out_channels, H, W).to(x.dtype)

    def forward_nhwc4(self, x4, threshold=0.5, want=("logits",)):
        """x4: normalised input [B,H,W,4], bf16 (fp32 when b200_precision == "fp32"). Returns (logits, probs, mask) with None
        for outputs not in `want`; shapes [B,H,W] for out_channels == 1 (the reference's case), [B,out_channels,H,W] otherwise."""
        self._check_input(x4, "input")
        B, H, W, _ = x4.shape
        dev = x4.device
        eng = self._engine(dev, H, W, B)
        want_dtype = torch.float32 if eng.precision == "fp32" else torch.bfloat16
        if x4.dtype != want_dtype or not x4.is_contiguous():
            raise ValueError(f"forward_nhwc4 with b200_precision={eng.precision!r} takes a contiguous {want_dtype} [B,H,W,4] tensor, got {x4.dtype}")
        self._last_engine = eng   # (a caller that captures this call in a CUDA graph keeps the reference: the graph replays into eng's buffers)
        shp = (B, H, W) if self.out_channels == 1 else (B, self.out_channels, H, W)
        logits = torch.empty(shp, dtype=torch.float32, device=dev) if "logits" in want else None
        probs = torch.empty(shp, dtype=torch.float32, device=dev) if "probs" in want else None
        mask = torch.empty(shp, dtype=torch.uint8, device=dev) if "mask" in want else None
        st = torch.cuda.current_stream().cuda_stream
        for b0 in range(0, B, eng.cap):
            n = min(eng.cap, B - b0)
            check(lib.unet_b200_forward(
                eng.handle, x4[b0:b0 + n].data_ptr(), n,
                None if logits is None else logits[b0:].data_ptr(),
                None if probs is None else probs[b0:].data_ptr(),
                None if mask is None else mask[b0:].data_ptr(), float(threshold), st))
            self.gpu_launches += eng.launches
        return logits, probs, mask

    @torch.no_grad()
    def predict_mask(self, frames_u8, threshold=0.5, swap_rb=False, size=(224, 224), want=("mask",)):
        """Fused inference pipeline on device-resident uint8 frames [B,Hs,Ws,3]:
        cv2-exact resize + normalise (src/unet.py:24-42, README.md:3110-3111) -> U-Net -> sigmoid ->
        (p > threshold)*255 (src/unet.py:63-67). Returns (logits, probs, mask) like forward_nhwc4."""
        self._check_input(frames_u8, "frames")
        if frames_u8.dim() != 4 or frames_u8.shape[3] != 3 or frames_u8.dtype != torch.uint8 or not frames_u8.is_contiguous():
            raise ValueError(f"expected contiguous uint8 frames [B,Hs,Ws,3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        if self.in_channels != 3:
            raise ValueError(f"predict_mask feeds 3-channel frames; this model has in_channels={self.in_channels}")
        B, Hs, Ws, _ = frames_u8.shape
        fp32 = self.b200_precision == "fp32"
        x4 = torch.empty(B, size[0], size[1], 4, dtype=torch.float32 if fp32 else torch.bfloat16, device=frames_u8.device)
        st = torch.cuda.current_stream().cuda_stream
        pre = lib.unet_b200_preprocess_u8_f32 if fp32 else lib.unet_b200_preprocess_u8
        check(pre(frames_u8.data_ptr(), B, Hs, Ws, Ws * 3, Hs * Ws * 3, size[0], size[1], int(swap_rb), f3(MEAN_255), f3(STD_255),
                  x4.data_ptr(), None, st))
        self.gpu_launches += 1
        return self.forward_nhwc4(x4, threshold=threshold, want=want)

    # ------------------------------------------------------------------ host-buffer entry (executor path)
    def infer_host(self, frames_host, threshold=0.5, swap_rb=False, size=(224, 224), mask_out=None, probs_out=None,
                   logits_out=None):
        """Reference-facing call with HOST buffers (RKNN_model_container.run contract): uint8 frames
        [B,Hs,Ws,3] on the host (pinned for full PCIe speed) -> host outputs. H2D copy, preprocess, U-Net,
        mask and D2H copy all go through unet_b200_infer_u8_host; returns after the results are on the host."""
        if (frames_host.is_cuda or frames_host.dtype != torch.uint8 or not frames_host.is_contiguous() or frames_host.dim() != 4
                or frames_host.shape[3] != 3):
            raise ValueError("frames_host must be a contiguous uint8 CPU tensor [B,Hs,Ws,3]")
        B, Hs, Ws, _ = frames_host.shape
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("UNet (B200): model parameters are not on a CUDA device (no CPU fallback)")
        eng = self._engine(dev, size[0], size[1], B)
        if lib.unet_b200_plan_host_pieces(eng.handle) == 0:
            eng = self._engine(dev, size[0], size[1], B, pipelined="halved")    # (pass-granular pipeline: two passes at least)
        skey = (Hs, Ws)
        if getattr(eng, "staging_key", None) != skey:
            eng.staging = torch.empty(lib.unet_b200_infer_stream_staging_bytes(eng.handle, Hs, Ws), dtype=torch.uint8, device=dev)
            eng.staging_key = skey
        st = torch.cuda.current_stream().cuda_stream
        # one call for the whole batch: passes of eng.cap frames; the copies are pipelined with the kernels piece by piece
        check(lib.unet_b200_infer_u8_host_stream(
            eng.handle, eng.staging.data_ptr(), frames_host.data_ptr(), B, Hs, Ws, int(swap_rb), f3(MEAN_255), f3(STD_255),
            float(threshold), None if logits_out is None else logits_out.data_ptr(),
            None if probs_out is None else probs_out.data_ptr(), None if mask_out is None else mask_out.data_ptr(), st))
        self.gpu_launches += lib.unet_b200_infer_stream_launches(eng.handle, B, Hs, Ws)
        return mask_out, probs_out, logits_out

    def profile_layers(self, x4):
        """Per-kernel durations (ms, CUDA events on the current stream) of one pass over x4 [B<=chunk,H,W,4],
        with each kernel's shape info and algorithmic FLOPs. Used by bench.py for the roofline line."""
        B, H, W, _ = x4.shape
        eng = self._engine(x4.device, H, W, B)
        if B > eng.cap:
            raise ValueError(f"profile_layers takes at most one chunk ({eng.cap} frames)")
        if eng.precision == "fp32":
            raise ValueError("profile_layers reports the bf16 plan's kernels; set b200_precision = 'bf16'")
        n = lib.unet_b200_plan_num_layers(eng.handle)
        ms = (C.c_float * n)()
        mask = torch.empty(B, H, W, dtype=torch.uint8, device=x4.device)
        check(lib.unet_b200_forward_profile(eng.handle, x4.data_ptr(), B, None, None, mask.data_ptr(), 0.5,
                                            torch.cuda.current_stream().cuda_stream, ms, n))
        rows = []
        kinds = {0: "stem", 1: "conv3x3", 2: "convT2x2", 3: "head"}
        for i in range(n):
            info = (C.c_int * 8)()
            check(lib.unet_b200_plan_layer_info(eng.handle, i, info))
            kind, h, w, cin, cout, taps, block_n, pool = list(info)
            mult = 4 if kind == 2 else 1
            cin_eff = self.in_channels if kind == 0 else cin
            flops = 2.0 * B * h * w * cout * mult * taps * cin_eff
            rows.append({"kind": kinds[kind], "H": h, "W": w, "Cin": cin, "Cout": cout, "block_n": block_n,
                         "fused_pool": bool(pool & 1), "halo": bool(pool & 2), "fused_head": bool(pool & 4), "ms": float(ms[i]), "flops": flops})
        return rows
