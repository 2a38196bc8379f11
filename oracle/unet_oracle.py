"""CPU oracle for the U-Net hot path of masktrump19-sudo/unet-lane-detection.

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does (it fails loudly when
its CUDA library is missing instead).

Parity status: the reference ships NO golden vectors or tests for this path, so
the oracle is pinned against (a) the structural facts the reference states - 31,037,633 parameters
(README.md:2288), output shape [1,1,224,224] (README.md:1489-1490) - and (b) outputs of the
reference's own `class UNet` listing executed verbatim in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz). cv2.resize parity of the preprocess is pinned
against cv2 itself on the reference's sample images (same script).

Each function cites the reference lines it restates (paths relative to the reference tree).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

IMAGENET_MEAN_255 = (123.675, 116.28, 103.53)   # README.md:3110 (RKNN config mean_values)
IMAGENET_STD_255 = (58.395, 57.12, 57.375)      # README.md:3111 (RKNN config std_values)
BN_EPS = 1e-5


# --------------------------------------------------------------------------------------------------
# Model: restates README.md:1421-1481 (class UNet). Attribute names and registration order are part
# of the contract (state_dict keys), so they are kept; the body is restated.
# --------------------------------------------------------------------------------------------------
def double_conv(cin: int, cout: int) -> nn.Sequential:
    """README.md:1449-1458: (Conv3x3 pad 1 no bias -> BatchNorm2d -> ReLU) x 2."""
    layers = []
    for a, b in ((cin, cout), (cout, cout)):
        layers += [nn.Conv2d(a, b, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(b), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class UNetOracle(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 1, features=(64, 128, 256, 512)):
        super().__init__()
        features = list(features)
        # registration order of README.md:1427-1447: encoder list, decoder list, pool, bottleneck, output
        self.encoder_blocks = nn.ModuleList()
        self.decoder_blocks = nn.ModuleList()
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        c = in_channels
        for f in features:
            self.encoder_blocks.append(double_conv(c, f))
            c = f
        self.bottleneck = double_conv(features[-1], 2 * features[-1])
        for f in features[::-1]:
            self.decoder_blocks.append(nn.ConvTranspose2d(2 * f, f, kernel_size=2, stride=2))
            self.decoder_blocks.append(double_conv(2 * f, f))
        self.output = nn.Conv2d(features[0], out_channels, kernel_size=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """README.md:1460-1481. Returns logits (no sigmoid)."""
        kept = []
        for block in self.encoder_blocks:
            x = block(x)
            kept.append(x)
            x = self.pool(x)
        x = self.bottleneck(x)
        for level in range(len(self.decoder_blocks) // 2):
            up = self.decoder_blocks[2 * level](x)
            skip = kept[-1 - level]
            x = self.decoder_blocks[2 * level + 1](torch.cat([skip, up], dim=1))  # skip first (README.md:1478)
        return self.output(x)


def randomize_bn_(model: nn.Module, seed: int = 1) -> None:
    """make BN non-trivial so a wrong fold is visible.
    gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            with torch.no_grad():
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)


def scale_head_(model: nn.Module, gain: float) -> None:
    """Give the logits a realistic scale (random init leaves |logit| ~ 1e-2)."""
    with torch.no_grad():
        model.output.weight.mul_(gain)


# --------------------------------------------------------------------------------------------------
# Loss: restates README.md:1855-1893 (BCEDiceLoss).
# --------------------------------------------------------------------------------------------------
class BCEDiceLossOracle(nn.Module):
    def __init__(self, bce_weight=0.5, dice_weight=0.5, pos_weight=None, smooth=1e-6):
        super().__init__()
        self.bce_weight, self.dice_weight, self.smooth = bce_weight, dice_weight, smooth
        self.bce = nn.BCEWithLogitsLoss(pos_weight=pos_weight)

    def forward(self, pred, target):
        target = target.float()
        bce = self.bce(pred, target)
        p = torch.sigmoid(pred).reshape(-1)
        t = target.reshape(-1)
        dice = 1 - (2.0 * (p * t).sum() + self.smooth) / (p.sum() + t.sum() + self.smooth)
        return self.bce_weight * bce + self.dice_weight * dice, bce, dice


def compute_dice_oracle(pred_mask: torch.Tensor, target: torch.Tensor, smooth: float = 1e-6) -> float:
    """README.md:2115-2120 (validation Dice on thresholded masks)."""
    p = pred_mask.reshape(-1).float()
    t = target.reshape(-1).float()
    return float((2.0 * (p * t).sum() + smooth) / (p.sum() + t.sum() + smooth))


# --------------------------------------------------------------------------------------------------
# Pre / post-processing: restates src/unet.py:24-72.
# --------------------------------------------------------------------------------------------------
def _linear_taps(dn: int, sn: int, clamp_weights: bool = True):
    """cv2 INTER_LINEAR tap positions and 11-bit fixed-point weights for one axis (uint8 path, resize.cpp).
    Horizontal axis (clamp_weights=True): a tap that falls outside the row is folded into its neighbour (fx = 0).
    Vertical axis (clamp_weights=False): cv2 keeps the weights and only clips the ROW INDICES, so at the top / bottom
    border both taps read the same row with weights (1-fy, fy) - two truncations instead of one."""
    scale = sn / dn
    f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = f - s
    if clamp_weights:
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= sn - 1
        f[hi], s[hi] = 0.0, sn - 1
    w1 = np.rint(f * np.float32(2048)).astype(np.int32)
    w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), w0, w1


def resize_bilinear_u8(img: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """Bit-exact numpy model of cv2.resize(img, (dw, dh)) for uint8 HxW[xC], default INTER_LINEAR (src/unet.py:33, :70),
    verified against cv2 4.13 for down- AND up-scaling (tests/golden/make_golden.py). Special cases of cv::resize:
    equal sizes copy; an exact 2x2 decimation is computed as INTER_AREA ((a+b+c+d+2)>>2)."""
    if img.ndim == 2:
        return resize_bilinear_u8(img[:, :, None], dh, dw)[:, :, 0]
    sh, sw = img.shape[:2]
    if (sh, sw) == (dh, dw):
        return img.copy()
    s = img.astype(np.int32)
    if sh == 2 * dh and sw == 2 * dw:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, ax0, ax1 = _linear_taps(dw, sw, True)
    y0, y1, by0, by1 = _linear_taps(dh, sh, False)
    horiz = s[:, x0, :] * ax0[None, :, None] + s[:, x1, :] * ax1[None, :, None]
    top, bot = horiz[y0], horiz[y1]
    out = (((by0[:, None, None] * (top >> 4)) >> 16) + ((by1[:, None, None] * (bot >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def invert3x3(m: np.ndarray) -> np.ndarray:
    """cv::invert of a 3x3 double matrix (closed form with the reciprocal determinant) - the inverse map that
    cv2.warpPerspective builds from its argument. Bit-equal to cv2.invert (make_golden.py)."""
    S = np.asarray(m, dtype=np.float64)
    det = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0])
           + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    d = 1.0 / det
    t = np.empty(9, dtype=np.float64)
    t[0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    t[1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    t[2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    t[3] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    t[4] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    t[5] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    t[6] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    t[7] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    t[8] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t.reshape(3, 3)


def warp_perspective_u8(src: np.ndarray, m_inv: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """Bit-exact numpy model of cv2.warpPerspective(src, M, (dw, dh)) (src/unet_ros_node.py:300-301: INTER_LINEAR,
    BORDER_CONSTANT 0) given m_inv = invert3x3(M). Follows imgwarp.cpp: destination coordinates are evaluated in double
    per 64-pixel column block (X0 + M0*x1)*W with W = 32/(W0 + M6*x1), rounded to 1/32 pixel; the four taps are blended
    with 15-bit integer weights (32-ax)(32-ay)*32 ... and rounded with +2^14 >> 15; taps outside the image are 0."""
    Hs, Ws = src.shape[:2]
    m = np.asarray(m_inv, dtype=np.float64).reshape(-1)
    yd = np.arange(dh, dtype=np.float64)[:, None]
    xd = np.arange(dw)[None, :]
    xb = (xd // 64) * 64
    x1 = (xd - xb).astype(np.float64)
    xb = xb.astype(np.float64)
    X0 = m[0] * xb + m[1] * yd + m[2]
    Y0 = m[3] * xb + m[4] * yd + m[5]
    W0 = m[6] * xb + m[7] * yd + m[8]
    W = W0 + m[6] * x1
    with np.errstate(divide="ignore"):
        W = np.where(W != 0, 32.0 / W, 0.0)
    fX = np.maximum(-2147483648.0, np.minimum(2147483647.0, (X0 + m[0] * x1) * W))
    fY = np.maximum(-2147483648.0, np.minimum(2147483647.0, (Y0 + m[3] * x1) * W))
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    ax, ay = X & 31, Y & 31
    s = src.astype(np.int64)
    if s.ndim == 2:
        s = s[:, :, None]

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < Ws) & (yy >= 0) & (yy < Hs)
        return s[np.clip(yy, 0, Hs - 1), np.clip(xx, 0, Ws - 1)] * inside[..., None]

    acc = (tap(sy, sx) * ((32 - ax) * (32 - ay) * 32)[..., None] + tap(sy, sx + 1) * (ax * (32 - ay) * 32)[..., None]
           + tap(sy + 1, sx) * ((32 - ax) * ay * 32)[..., None] + tap(sy + 1, sx + 1) * (ax * ay * 32)[..., None])
    out = np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return out.reshape(dh, dw, *src.shape[2:])


def ipm_preprocess_oracle(bgr_u8: np.ndarray, m_inv: np.ndarray, warp_size=(1055, 685), size=(224, 224)):
    """src/unet_ros_node.py:297-313 up to the network input: warpPerspective to warp_size (w, h) -> resize to the same
    size (a copy) -> BGR2RGB -> RKNNLaneInference.preprocess_image (resize to the model input, src/unet.py:24-42).
    Returns (uint8 NHWC [1,H,W,3] RGB, (h, w) of the warped image = the size the mask is resized back to)."""
    warped = warp_perspective_u8(bgr_u8, m_inv, warp_size[0], warp_size[1])
    rgb = np.ascontiguousarray(warped[:, :, ::-1])
    return preprocess_oracle(rgb, size)


def preprocess_oracle(image_u8: np.ndarray, size=(224, 224), swap_rb: bool = False):
    """src/unet.py:24-42: resize to the model input, keep uint8, add the batch dim -> (1,H,W,3).
    swap_rb models the caller's BGR->RGB (src/unet_ros_node.py:310). Returns (uint8 NHWC, original_shape)."""
    original_shape = image_u8.shape[:2]
    img = resize_bilinear_u8(image_u8, size[0], size[1])
    if swap_rb:
        img = img[:, :, ::-1]
    return np.ascontiguousarray(img)[None], original_shape


def normalize_oracle(frames_u8_nhwc: np.ndarray) -> np.ndarray:
    """In-graph normalisation of the deployed model (README.md:3110-3111): (x - mean) / std -> NCHW fp32."""
    x = frames_u8_nhwc.astype(np.float32)
    x = (x - np.asarray(IMAGENET_MEAN_255, np.float32)) / np.asarray(IMAGENET_STD_255, np.float32)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def postprocess_oracle(output, original_shape, threshold: float = 0.5) -> np.ndarray:
    """src/unet.py:44-72: take [0,0], int8->f32, sigmoid iff values leave [0,1], strict '>' threshold,
    *255 uint8, resize back to the source size."""
    mask = output[0] if isinstance(output, (list, tuple)) and len(output) > 0 else output
    mask = np.asarray(mask)
    if mask.ndim == 4:
        mask = mask[0, 0]
    elif mask.ndim == 3:
        mask = mask[0]
    if mask.dtype == np.int8:
        mask = mask.astype(np.float32)
    if mask.max() > 1.0 or mask.min() < 0.0:
        mask = 1 / (1 + np.exp(-mask))
    binary = (mask > threshold).astype(np.uint8) * 255
    return resize_bilinear_u8(binary, original_shape[0], original_shape[1])


# --------------------------------------------------------------------------------------------------
# Layer-level helpers used by the kernel parity tests.
# --------------------------------------------------------------------------------------------------
def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def fold_bn(conv_w: torch.Tensor, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm folded into the preceding bias-free conv: w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return conv_w * s[:, None, None, None], bn.bias - bn.running_mean * s


@torch.no_grad()
def forward_bf16_emulated(model: UNetOracle, x: torch.Tensor, return_feats: bool = False):
    """The same network evaluated with the B200 path's rounding points: folded weights rounded to
    bf16, activations rounded to bf16 at every layer boundary, fp32 accumulation, fp32 head.
    This is NOT the parity target (that is model(x) in fp32); it separates 'bf16 by design' error
    from kernel bugs in the per-layer tests."""
    feats = {}

    def block(seq, t, name):
        for k, (ci, bi) in enumerate(((0, 1), (3, 4))):
            w, b = fold_bn(seq[ci].weight, seq[bi])
            t = bf16_round(F.relu(F.conv2d(t, bf16_round(w), b, padding=1)))
            feats[f"{name}.{k}"] = t
        return t

    t = bf16_round(x)
    kept = []
    for i, enc in enumerate(model.encoder_blocks):
        t = block(enc, t, f"enc{i}")
        kept.append(t)
        t = F.max_pool2d(t, 2)
    t = block(model.bottleneck, t, "bott")
    for level in range(len(model.decoder_blocks) // 2):
        up = model.decoder_blocks[2 * level]
        t = bf16_round(F.conv_transpose2d(t, bf16_round(up.weight), up.bias, stride=2))
        feats[f"up{level}"] = t
        t = block(model.decoder_blocks[2 * level + 1], torch.cat([kept[-1 - level], t], dim=1), f"dec{level}")
    logits = F.conv2d(t, model.output.weight, model.output.bias)
    return (logits, feats) if return_feats else logits


class _RoundBoth(torch.autograd.Function):
    """bf16 rounding point for activations: the value is rounded forward, its gradient backward (both live in bf16)."""

    @staticmethod
    def forward(ctx, t):
        return bf16_round(t)

    @staticmethod
    def backward(ctx, g):
        return bf16_round(g)


class _RoundFwd(torch.autograd.Function):
    """bf16 rounding point for weights: the operand copy is bf16, the weight gradient stays fp32."""

    @staticmethod
    def forward(ctx, t):
        return bf16_round(t)

    @staticmethod
    def backward(ctx, g):
        return g


def forward_train_bf16_emulated(model: UNetOracle, x: torch.Tensor) -> torch.Tensor:
    """Training-mode forward (batch-statistics BatchNorm, README.md:2062) with the B200 path's rounding points,
    differentiable: conv operands, raw conv outputs y, activations a = relu(bn(y)) and every activation gradient are
    rounded to bf16; statistics, accumulation, weight gradients and the head stay fp32. Like forward_bf16_emulated this
    is NOT the parity target (model(x) in fp32 is); it separates rounding-by-design (ReLU / max-pool decisions that flip
    on near-ties) from kernel bugs in the gradient tests. Running statistics of `model` are updated as a side effect."""
    r, rw = _RoundBoth.apply, _RoundFwd.apply

    def block(seq, t):
        for ci, bi in ((0, 1), (3, 4)):
            bn = seq[bi]
            y = r(F.conv2d(t, rw(seq[ci].weight), None, padding=1))
            t = r(F.relu(F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)))
            with torch.no_grad():
                bn.num_batches_tracked += 1
        return t

    t = r(x)
    kept = []
    for enc in model.encoder_blocks:
        t = block(enc, t)
        kept.append(t)
        t = F.max_pool2d(t, 2)
    t = block(model.bottleneck, t)
    for level in range(len(model.decoder_blocks) // 2):
        up = model.decoder_blocks[2 * level]
        t = r(F.conv_transpose2d(t, rw(up.weight), up.bias, stride=2))
        t = block(model.decoder_blocks[2 * level + 1], torch.cat([kept[-1 - level], t], dim=1))
    return F.conv2d(t, model.output.weight, model.output.bias)


def mask_agreement(logits_a: torch.Tensor, logits_b: torch.Tensor, threshold: float = 0.5, band: float = 0.0):
    """Fraction of pixels whose thresholded masks agree; pixels with |logit_b - logit(thr)| < band excluded."""
    z = float(np.log(threshold / (1 - threshold)))
    a, b = logits_a.reshape(-1) > z, logits_b.reshape(-1) > z
    keep = (logits_b.reshape(-1) - z).abs() >= band
    if keep.sum() == 0:
        return 1.0
    return float((a[keep] == b[keep]).float().mean())
