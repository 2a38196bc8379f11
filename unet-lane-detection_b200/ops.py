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


def pack_stem(w, bn=None, fp32=False):
    """fp32=False: weights rounded through bf16 (the bf16 path's stem); fp32=True: kept in fp32 (split-precision path)."""
    _req(w, torch.float32, "w")
    cout, cin = w.shape[:2]
    ws = torch.empty(9, 4, cout, dtype=torch.float32, device=w.device)
    bias = torch.empty(cout, dtype=torch.float32, device=w.device)
    fn = lib.unet_b200_pack_stem_fp32 if fp32 else lib.unet_b200_pack_stem
    g, b, m, v, eps = bn if bn is not None else (None, None, None, None, 0.0)
    check(fn(w.data_ptr(), _p(g), _p(b), _p(m), _p(v), eps, cout, cin, ws.data_ptr(), bias.data_ptr(), _stream()))
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


# ------------------------------------------------------------------------------------------------
# Training ops (include/unet_b200.h "single training ops"): thin wrappers used by the parity tests.
# ------------------------------------------------------------------------------------------------
def pack_conv3x3_dgrad(w):
    """w fp32 [Cout,Cin,3,3] -> wd bf16 [Cin,9,Cout] (rotated 180 degrees, in/out swapped)."""
    _req(w, torch.float32, "w")
    cout, cin = w.shape[:2]
    wd = torch.empty(cin, 9, cout, dtype=torch.bfloat16, device=w.device)
    check(lib.unet_b200_pack_conv3x3_dgrad(w.data_ptr(), cout, cin, wd.data_ptr(), _stream()))
    return wd


def conv3x3_dgrad(dy, wd):
    """Input gradient of a bias-free 3x3 conv: dy bf16 [B,H,W,Cout], wd from pack_conv3x3_dgrad -> dx bf16 [B,H,W,Cin]."""
    zero = torch.zeros(wd.shape[0], dtype=torch.float32, device=dy.device)
    return conv3x3(dy, wd, zero, relu=False)


def conv3x3_wgrad(x0, dy, x1=None):
    """dW fp32 [Cout, C0+C1, 3, 3] for x = cat(x0, x1) and dy."""
    _req(x0, torch.bfloat16, "x0")
    _req(dy, torch.bfloat16, "dy")
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else _req(x1, torch.bfloat16, "x1").shape[3]
    cout = dy.shape[3]
    dw = torch.zeros(cout, C0 + C1, 3, 3, dtype=torch.float32, device=x0.device)
    check(lib.unet_b200_conv3x3_wgrad(x0.data_ptr(), C0, _p(x1), C1, dy.data_ptr(), B, H, W, cout, dw.data_ptr(), _stream()))
    return dw


def stem_wgrad(x4, dy, cin):
    _req(x4, torch.bfloat16, "x4")
    _req(dy, torch.bfloat16, "dy")
    B, H, W, _ = x4.shape
    cout = dy.shape[3]
    dw = torch.zeros(cout, cin, 3, 3, dtype=torch.float32, device=x4.device)
    check(lib.unet_b200_stem_wgrad(x4.data_ptr(), dy.data_ptr(), B, H, W, cin, cout, dw.data_ptr(), _stream()))
    return dw


def pack_convT2x2_dgrad(w):
    _req(w, torch.float32, "w")
    cin, f = w.shape[:2]
    wd = torch.empty(cin, 4 * f, dtype=torch.bfloat16, device=w.device)
    check(lib.unet_b200_pack_convT2x2_dgrad(w.data_ptr(), cin, f, wd.data_ptr(), _stream()))
    return wd


def _dup_view(dup, f):
    """dup: bf16 [B,2H,2W,P] holding the ConvT output gradient in its LAST f channels (P >= f)."""
    _req(dup, torch.bfloat16, "dup")
    pitch = dup.shape[3]
    return dup.data_ptr() + (pitch - f) * 2, pitch


def convT2x2_wgrad(x, dup, f):
    """x bf16 [B,H,W,Cin]; dup bf16 [B,2H,2W,P] (gradient in the last f channels) -> (dW fp32 [Cin,f,2,2], dbias fp32 [f])."""
    _req(x, torch.bfloat16, "x")
    B, H, W, cin = x.shape
    ptr, pitch = _dup_view(dup, f)
    dw = torch.zeros(cin, f, 2, 2, dtype=torch.float32, device=x.device)
    db = torch.zeros(f, dtype=torch.float32, device=x.device)
    check(lib.unet_b200_convT2x2_wgrad(x.data_ptr(), cin, ptr, pitch, B, H, W, f, dw.data_ptr(), db.data_ptr(), _stream()))
    return dw, db


def convT2x2_dgrad(dup, wd, f):
    ptr, pitch = _dup_view(dup, f)
    B, H2, W2, _ = dup.shape
    cin = wd.shape[0]
    dx = torch.empty(B, H2 // 2, W2 // 2, cin, dtype=torch.bfloat16, device=dup.device)
    check(lib.unet_b200_convT2x2_dgrad(ptr, pitch, wd.data_ptr(), B, H2 // 2, W2 // 2, cin, f, dx.data_ptr(), _stream()))
    return dx


def bn_relu_train_fwd(y, gamma, beta, eps=1e-5, momentum=0.1, running_mean=None, running_var=None, pool=False):
    """y bf16 [B,H,W,C] -> (a, pooled or None, stats fp32 [4,C] = mean, invstd, scale, shift)."""
    _req(y, torch.bfloat16, "y")
    B, H, W, C = y.shape
    a = torch.empty_like(y)
    p = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=y.device) if pool else None
    stats = torch.empty(4, C, dtype=torch.float32, device=y.device)
    scratch = torch.empty(2 * C, dtype=torch.float64, device=y.device)
    check(lib.unet_b200_bn_relu_train_fwd(y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, H, W, C, float(eps),
                                          float(momentum), _p(running_mean), _p(running_var), a.data_ptr(), _p(p),
                                          stats.data_ptr(), scratch.data_ptr(), _stream()))
    return a, p, stats


def bn_relu_bwd(dA, y, stats):
    """dA bf16 [B,H,W,C] (not modified) -> (dY bf16, dgamma fp32 [C], dbeta fp32 [C])."""
    _req(dA, torch.bfloat16, "dA")
    _req(y, torch.bfloat16, "y")
    B, H, W, C = y.shape
    g = dA.clone()
    dgamma = torch.empty(C, dtype=torch.float32, device=y.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=y.device)
    check(lib.unet_b200_bn_relu_bwd(g.data_ptr(), y.data_ptr(), stats.data_ptr(), B, H, W, C, dgamma.data_ptr(),
                                    dbeta.data_ptr(), _stream()))
    return g, dgamma, dbeta


def maxpool2x2_bwd(a, dP, dskip=None):
    """a bf16 [B,H,W,C], dP bf16 [B,H/2,W/2,C], dskip bf16 [B,H,W,P>=C] (its FIRST C channels are added) -> dA."""
    _req(a, torch.bfloat16, "a")
    _req(dP, torch.bfloat16, "dP")
    B, H, W, C = a.shape
    dA = torch.empty_like(a)
    pitch = C if dskip is None else _req(dskip, torch.bfloat16, "dskip").shape[3]
    check(lib.unet_b200_maxpool2x2_bwd(a.data_ptr(), dP.data_ptr(), _p(dskip), pitch, B, H, W, C, dA.data_ptr(), _stream()))
    return dA


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
               grad_scale=1.0):
    """In-place torch.optim.AdamW step on flat fp32 tensors."""
    for t, n in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, n)
    check(lib.unet_b200_adamw_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), params.numel(),
                                   float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                   float(grad_scale), _stream()))


# ------------------------------------------------------------------------------------------------
# Camera-side steps of the ROS node (src/unet_ros_node.py:297-313) and the mask up-resize (src/unet.py:70).
# ------------------------------------------------------------------------------------------------
def invert3x3(m):
    """cv::invert of a 3x3 double matrix (closed form, reciprocal determinant): the inverse map cv2.warpPerspective
    derives from its matrix argument. Host-side, float64, bit-equal to cv2.invert."""
    import numpy as np
    S = np.asarray(m, dtype=np.float64)
    det = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0])
           + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    d = 1.0 / det
    t = [(S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d, (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d,
         (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d, (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d,
         (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d, (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d,
         (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d, (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d,
         (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d]
    return np.asarray(t, dtype=np.float64).reshape(3, 3)


def _m9(matrix):
    import ctypes as C
    inv = invert3x3(matrix).reshape(-1)
    return (C.c_double * 9)(*[float(v) for v in inv])


def preprocess_warp_u8(frames, matrix, warp_size=(1055, 685), size=(224, 224), swap_rb=True, mean=MEAN_255, std=STD_255,
                       return_resized=False):
    """frames uint8 [B,Hs,Ws,3] BGR (CUDA); matrix: the 3x3 perspective matrix handed to cv2.warpPerspective;
    warp_size (w, h) as in cv2. Returns NHWC4 bf16 [B,H,W,4] (+ the resized uint8 RGB frames)."""
    _req(frames, torch.uint8, "frames")
    B, Hs, Ws, c = frames.shape
    if c != 3:
        raise ValueError("frames must be [B,Hs,Ws,3]")
    H, W = size
    y = torch.empty(B, H, W, 4, dtype=torch.bfloat16, device=frames.device)
    r = torch.empty(B, H, W, 3, dtype=torch.uint8, device=frames.device) if return_resized else None
    check(lib.unet_b200_preprocess_warp_u8(frames.data_ptr(), B, Hs, Ws, Ws * 3, Hs * Ws * 3, _m9(matrix), warp_size[1],
                                           warp_size[0], H, W, int(swap_rb), f3(mean), f3(std), y.data_ptr(), _p(r), None,
                                           _stream()))
    return (y, r) if return_resized else y


def warp_perspective_u8(frames, matrix, warp_size=(1055, 685)):
    """cv2.warpPerspective(frame, matrix, warp_size) for a batch of uint8 [B,Hs,Ws,3] frames (bit-exact)."""
    _req(frames, torch.uint8, "frames")
    B, Hs, Ws, _ = frames.shape
    out = torch.empty(B, warp_size[1], warp_size[0], 3, dtype=torch.uint8, device=frames.device)
    check(lib.unet_b200_preprocess_warp_u8(frames.data_ptr(), B, Hs, Ws, Ws * 3, Hs * Ws * 3, _m9(matrix), warp_size[1],
                                           warp_size[0], 1, 1, 0, f3(MEAN_255), f3(STD_255), None, None, out.data_ptr(),
                                           _stream()))
    return out


def resize_gray_u8(masks, size):
    """cv2.resize(mask, (w, h)) for uint8 [B,Hs,Ws] masks; size = (h, w). Bit-exact INTER_LINEAR."""
    _req(masks, torch.uint8, "masks")
    B, Hs, Ws = masks.shape
    out = torch.empty(B, size[0], size[1], dtype=torch.uint8, device=masks.device)
    check(lib.unet_b200_resize_gray_u8(masks.data_ptr(), B, Hs, Ws, out.data_ptr(), size[0], size[1], _stream()))
    return out


# ---------------------------------------------------------------- split-precision ("fp32-class") layers
# Tensors are bf16 [B,H,W,2C] = [hi C | lo C] with value = hi + lo (16 mantissa bits); every product is evaluated as
# hi*hi + lo*hi + hi*lo in the fp32 tensor-core accumulator (csrc/aux_kernels.cuh, csrc/conv_umma.cuh `split`).
def pack_conv3x3_split(w, bn=None, c0=None):
    """w fp32 [Cout,Cin,3,3] (+ eval BN); c0 = channels of the first input tensor (skip) when the conv reads a concat.
    -> (wp bf16 [Cout, 9, 3*Cin], bias fp32 [Cout])."""
    _req(w, torch.float32, "w")
    cout, cin = w.shape[:2]
    c0 = cin if c0 is None else c0
    wp = torch.empty(cout, 9, 3 * cin, dtype=torch.bfloat16, device=w.device)
    bias = torch.empty(cout, dtype=torch.float32, device=w.device)
    g, b, m, v, eps = bn if bn is not None else (None, None, None, None, 0.0)
    check(lib.unet_b200_pack_conv3x3_split(w.data_ptr(), _p(g), _p(b), _p(m), _p(v), eps, cout, c0, cin - c0, wp.data_ptr(),
                                           bias.data_ptr(), _stream()))
    return wp, bias


def conv3x3_split(x0, wp, bias, x1=None, relu=True):
    _req(x0, torch.bfloat16, "x0")
    B, H, W, C0 = x0.shape
    C1 = 0
    if x1 is not None:
        _req(x1, torch.bfloat16, "x1")
        C1 = x1.shape[3]
    cout = wp.shape[0]
    y = torch.empty(B, H, W, 2 * cout, dtype=torch.bfloat16, device=x0.device)
    check(lib.unet_b200_conv3x3_split(x0.data_ptr(), C0 // 2, _p(x1), C1 // 2, wp.data_ptr(), bias.data_ptr(), B, H, W, cout,
                                      int(relu), y.data_ptr(), _stream()))
    return y


def pack_convT2x2_split(w):
    """w fp32 [Cin,f,2,2] -> wp bf16 [4f, 3*Cin]."""
    _req(w, torch.float32, "w")
    cin, f = w.shape[:2]
    wp = torch.empty(4 * f, 3 * cin, dtype=torch.bfloat16, device=w.device)
    check(lib.unet_b200_pack_convT2x2_split(w.data_ptr(), cin, f, wp.data_ptr(), _stream()))
    return wp


def convT2x2_split(x, wp, bias):
    _req(x, torch.bfloat16, "x")
    B, H, W, c2 = x.shape
    f = wp.shape[0] // 4
    y = torch.empty(B, 2 * H, 2 * W, 2 * f, dtype=torch.bfloat16, device=x.device)
    check(lib.unet_b200_convT2x2_split(x.data_ptr(), c2 // 2, wp.data_ptr(), bias.data_ptr(), B, H, W, f, y.data_ptr(), _stream()))
    return y


def stem_conv_split(x_nchw, ws, bias, relu=True):
    """x fp32 NCHW [B,Cin<=4,H,W] (read as is, no bf16 rounding of the image), ws/bias from pack_stem -> [B,H,W,2*Cout]."""
    _req(x_nchw, torch.float32, "x")
    B, cin, H, W = x_nchw.shape
    cout = ws.shape[2]
    y = torch.empty(B, H, W, 2 * cout, dtype=torch.bfloat16, device=x_nchw.device)
    check(lib.unet_b200_stem_conv_split(x_nchw.data_ptr(), ws.data_ptr(), bias.data_ptr(), B, H, W, cin, cout, int(relu),
                                        y.data_ptr(), _stream()))
    return y


def maxpool2x2_split(x):
    _req(x, torch.bfloat16, "x")
    B, H, W, c2 = x.shape
    y = torch.empty(B, H // 2, W // 2, c2, dtype=torch.bfloat16, device=x.device)
    check(lib.unet_b200_maxpool2x2_split(x.data_ptr(), B, H, W, c2 // 2, y.data_ptr(), _stream()))
    return y
