"""Thin torch-tensor wrappers over the single-layer C-ABI entry points (include/unet_b200.h).
torch is used for device memory and the current stream only."""
import torch

from ._lib import check, f3, lib

MEAN_255 = (123.675, 116.28, 103.53)  # reference README.md:3110
STD_255 = (58.395, 57.12, 57.375)     # reference README.md:3111


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _p(t):
    return None if t is None else t.data_ptr()


def pack_conv3x3(w, bn=None):
    """w fp32 [Cout,Cin,3,3] (+ optional eval BatchNorm2d) -> (wp bf16 [Cout,9,Cin], bias fp32 [Cout])."""
    _req(w, torch.float32, "w")
    cout, cin = w.shape[:2]
    wp = torch.empty(cout, 9, cin, dtype=torch.bfloat16, device=w.device)
    bias = torch.empty(cout, dtype=torch.float32, device=w.device)
    if bn is None:
        check(lib.unet_b200_pack_conv3x3(w.data_ptr(), None, None, None, None, 0.0, cout, cin, wp.data_ptr(),
                                         bias.data_ptr(), _stream()))
    else:
        g, b, m, v, eps = bn
        check(lib.unet_b200_pack_conv3x3(w.data_ptr(), g.data_ptr(), b.data_ptr(), m.data_ptr(), v.data_ptr(), eps,
                                         cout, cin, wp.data_ptr(), bias.data_ptr(), _stream()))
    return wp, bias


def pack_stem(w, bn=None):
    _req(w, torch.float32, "w")
    cout, cin = w.shape[:2]
    ws = torch.empty(9, 4, cout, dtype=torch.float32, device=w.device)
    bias = torch.empty(cout, dtype=torch.float32, device=w.device)
    if bn is None:
        check(lib.unet_b200_pack_stem(w.data_ptr(), None, None, None, None, 0.0, cout, cin, ws.data_ptr(),
                                      bias.data_ptr(), _stream()))
    else:
        g, b, m, v, eps = bn
        check(lib.unet_b200_pack_stem(w.data_ptr(), g.data_ptr(), b.data_ptr(), m.data_ptr(), v.data_ptr(), eps,
                                      cout, cin, ws.data_ptr(), bias.data_ptr(), _stream()))
    return ws, bias


def pack_stem_tc(w, bn=None):
    """Tensor-core stem weights: w fp32 [64,Cin<=4,3,3] (+ eval BN) -> (wp bf16 [64,64], bias fp32 [64])."""
    _req(w, torch.float32, "w")
    cout, cin = w.shape[:2]
    wp = torch.empty(cout, 64, dtype=torch.bfloat16, device=w.device)
    bias = torch.empty(cout, dtype=torch.float32, device=w.device)
    g, b, m, v, eps = bn if bn is not None else (None, None, None, None, 0.0)
    check(lib.unet_b200_pack_stem_tc(w.data_ptr(), _p(g), _p(b), _p(m), _p(v), eps, cout, cin, wp.data_ptr(),
                                     bias.data_ptr(), _stream()))
    return wp, bias


def stem_conv_tc(x4, wp, bias, relu=True):
    _req(x4, torch.bfloat16, "x4")
    B, H, W, _ = x4.shape
    y = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=x4.device)
    check(lib.unet_b200_stem_conv_tc(x4.data_ptr(), wp.data_ptr(), bias.data_ptr(), B, H, W, int(relu), y.data_ptr(),
                                     _stream()))
    return y


def pack_convT2x2(w):
    """w fp32 [Cin,f,2,2] -> wp bf16 [4f, Cin]."""
    _req(w, torch.float32, "w")
    cin, f = w.shape[:2]
    wp = torch.empty(4 * f, cin, dtype=torch.bfloat16, device=w.device)
    check(lib.unet_b200_pack_convT2x2(w.data_ptr(), cin, f, wp.data_ptr(), _stream()))
    return wp


def conv3x3(x0, wp, bias, x1=None, relu=True, pool=False):
    """x0/x1 bf16 NHWC; returns y (and pooled y when pool=True)."""
    _req(x0, torch.bfloat16, "x0")
    B, H, W, C0 = x0.shape
    C1 = 0
    if x1 is not None:
        _req(x1, torch.bfloat16, "x1")
        C1 = x1.shape[3]
    cout = wp.shape[0]
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=x0.device)
    yp = torch.empty(B, H // 2, W // 2, cout, dtype=torch.bfloat16, device=x0.device) if pool else None
    check(lib.unet_b200_conv3x3(x0.data_ptr(), C0, _p(x1), C1, wp.data_ptr(), bias.data_ptr(), B, H, W, cout,
                                int(relu), y.data_ptr(), _p(yp), _stream()))
    return (y, yp) if pool else y


def convT2x2(x, wp, bias):
    _req(x, torch.bfloat16, "x")
    B, H, W, cin = x.shape
    f = wp.shape[0] // 4
    y = torch.empty(B, 2 * H, 2 * W, f, dtype=torch.bfloat16, device=x.device)
    check(lib.unet_b200_convT2x2(x.data_ptr(), cin, wp.data_ptr(), bias.data_ptr(), B, H, W, f, y.data_ptr(), _stream()))
    return y


def stem_conv(x4, ws, bias, cin, relu=True):
    _req(x4, torch.bfloat16, "x4")
    B, H, W, _ = x4.shape
    cout = ws.shape[2]
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=x4.device)
    check(lib.unet_b200_stem_conv(x4.data_ptr(), ws.data_ptr(), bias.data_ptr(), B, H, W, cin, cout, int(relu),
                                  y.data_ptr(), _stream()))
    return y


def head(x, w, bias: float, threshold=0.5, want=("logits", "probs", "mask")):
    _req(x, torch.bfloat16, "x")
    B, H, W, C = x.shape
    dev = x.device
    logits = torch.empty(B, H, W, dtype=torch.float32, device=dev) if "logits" in want else None
    probs = torch.empty(B, H, W, dtype=torch.float32, device=dev) if "probs" in want else None
    mask = torch.empty(B, H, W, dtype=torch.uint8, device=dev) if "mask" in want else None
    check(lib.unet_b200_head(x.data_ptr(), w.data_ptr(), float(bias), B * H * W, C, _p(logits), _p(probs), _p(mask),
                             float(threshold), _stream()))
    return logits, probs, mask


def maxpool2x2(x):
    _req(x, torch.bfloat16, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=x.device)
    check(lib.unet_b200_maxpool2x2(x.data_ptr(), B, H, W, C, y.data_ptr(), _stream()))
    return y


def nchw_to_nhwc4(x):
    _req(x, torch.float32, "x")
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, 4, dtype=torch.bfloat16, device=x.device)
    check(lib.unet_b200_nchw_to_nhwc4(x.data_ptr(), B, C, H, W, y.data_ptr(), _stream()))
    return y


def preprocess_u8(frames, size=(224, 224), swap_rb=False, mean=MEAN_255, std=STD_255, return_resized=False):
    """frames uint8 [B,Hs,Ws,3] (CUDA) -> NHWC4 bf16 [B,H,W,4] normalised; optionally the resized uint8 frames."""
    _req(frames, torch.uint8, "frames")
    B, Hs, Ws, c = frames.shape
    if c != 3:
        raise ValueError("frames must be [B,Hs,Ws,3]")
    H, W = size
    y = torch.empty(B, H, W, 4, dtype=torch.bfloat16, device=frames.device)
    r = torch.empty(B, H, W, 3, dtype=torch.uint8, device=frames.device) if return_resized else None
    check(lib.unet_b200_preprocess_u8(frames.data_ptr(), B, Hs, Ws, Ws * 3, Hs * Ws * 3, H, W, int(swap_rb), f3(mean),
                                      f3(std), y.data_ptr(), _p(r), _stream()))
    return (y, r) if return_resized else y
