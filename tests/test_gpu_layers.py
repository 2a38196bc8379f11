"""GPU (B200): every kernel of the hot path, through the C ABI, against the CPU oracle on the same
seeded inputs. Tolerances: conv/convT outputs are bf16 -> at most 1 bf16 ulp of the value range plus
accumulation-order noise (2e-2 * max(1, |ref|max)); integer/uint8 work (resize, pool, mask) is bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
bf = O.bf16_round


@pytest.fixture(scope="module")
def U():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import unet_lane_detection_b200 as mod
    from unet_lane_detection_b200._lib import lib
    assert lib.unet_b200_device_ok() == 0, lib.unet_b200_last_error()
    return mod


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw(y):
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def close(got, ref, tol):
    d = (got.float().cpu() - ref.float().cpu()).abs().max().item()
    lim = tol * max(1.0, ref.abs().max().item())
    assert d <= lim, f"max|d|={d:.4e} > {lim:.4e}"


def conv_case(U, B, H, W, C0, C1, Cout, pool=False, relu=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    x0 = bf(torch.randn(B, C0, H, W, generator=g))
    x1 = bf(torch.randn(B, C1, H, W, generator=g)) if C1 else None
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=g) / (3.0 * (C0 + C1) ** 0.5)
    ga, be = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1
    mu, va = torch.randn(Cout, generator=g) * 0.1, torch.rand(Cout, generator=g) + 0.5
    wp, bias = U.pack_conv3x3(w.cuda(), (ga.cuda(), be.cuda(), mu.cuda(), va.cuda(), 1e-5))
    s = ga / torch.sqrt(va + 1e-5)
    ref = F.conv2d(torch.cat([x0, x1], 1) if C1 else x0, bf(w * s[:, None, None, None]), be - mu * s, padding=1)
    if relu:
        ref = F.relu(ref)
    out = U.conv3x3(nhwc(x0), wp, bias, x1=nhwc(x1) if C1 else None, relu=relu, pool=pool)
    y = out[0] if pool else out
    close(nchw(y), bf(ref), 2e-2)
    if pool:
        assert torch.equal(nchw(out[1]), F.max_pool2d(nchw(y), 2)), "fused 2x2 max-pool is not bit-exact"


# every 3x3 conv shape of the default network at a small batch
@pytest.mark.parametrize("B,H,C0,Cout", [(1, 224, 64, 64), (1, 112, 64, 128), (1, 112, 128, 128), (2, 56, 128, 256),
                                         (2, 56, 256, 256), (8, 28, 256, 512), (8, 28, 512, 512), (32, 14, 512, 1024),
                                         (32, 14, 1024, 1024)])
def test_conv3x3_network_shapes(U, B, H, C0, Cout):
    conv_case(U, B, H, H, C0, 0, Cout)


@pytest.mark.parametrize("B,H,W,f", [(1, 224, 224, 64), (1, 112, 112, 128), (2, 56, 56, 256), (8, 28, 28, 512)])
def test_conv3x3_concat_is_two_source_k_loop(U, B, H, W, f):
    """decoder conv0 = conv over cat([skip, up]) without materialising the concat (README.md:1478: skip first)."""
    conv_case(U, B, H, W, f, f, f)


@pytest.mark.parametrize("B,H,W,C", [(1, 224, 224, 64), (1, 112, 112, 128), (2, 56, 56, 256), (8, 28, 28, 512)])
def test_conv3x3_fused_pool(U, B, H, W, C):
    conv_case(U, B, H, W, C, 0, C, pool=True)


@pytest.mark.parametrize("B,H,W", [(1, 16, 16), (3, 14, 14), (5, 28, 28), (1, 60, 80), (1, 480, 640), (7, 8, 24)])
def test_conv3x3_ragged_batches_and_sizes(U, B, H, W):
    """Batch sizes that do not fill the pixel box, non-square and camera-resolution grids."""
    conv_case(U, B, H, W, 64, 0, 64, pool=(H % 2 == 0 and W % 2 == 0))


def test_conv3x3_no_relu_and_no_bn(U):
    conv_case(U, 1, 32, 32, 64, 0, 128, relu=False)
    x = bf(torch.randn(1, 64, 16, 16))
    w = torch.randn(64, 64, 3, 3) / 24
    wp, bias = U.pack_conv3x3(w.cuda())
    assert bias.abs().max().item() == 0.0
    close(nchw(U.conv3x3(nhwc(x), wp, bias, relu=False)), bf(F.conv2d(x, bf(w), padding=1)), 2e-2)


def test_conv3x3_linearity(U):
    """Size-independent property: without bias/ReLU the kernel is linear in its input (power-of-two scaling is exact)."""
    x = bf(torch.randn(2, 64, 56, 56))
    w = torch.randn(128, 64, 3, 3) / 24
    wp, bias = U.pack_conv3x3(w.cuda())
    y1 = U.conv3x3(nhwc(x), wp, bias, relu=False)
    y2 = U.conv3x3(nhwc(x * 4), wp, bias, relu=False)
    assert torch.equal(y2.float(), y1.float() * 4)


@pytest.mark.parametrize("B,H,Cin,f", [(32, 14, 1024, 512), (8, 28, 512, 256), (2, 56, 256, 128), (1, 112, 128, 64), (3, 14, 128, 64)])
def test_convT2x2(U, B, H, Cin, f):
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(B, Cin, H, H, generator=g))
    w = torch.randn(Cin, f, 2, 2, generator=g) / Cin ** 0.5
    b = torch.randn(f, generator=g) * 0.1
    y = U.convT2x2(nhwc(x), U.pack_convT2x2(w.cuda()), b.cuda())
    close(nchw(y), bf(F.conv_transpose2d(x, bf(w), b, stride=2)), 2e-2)


@pytest.mark.parametrize("cin,cout,H,W", [(3, 64, 224, 224), (3, 64, 32, 48), (1, 64, 16, 16), (4, 128, 24, 40), (3, 32, 16, 16)])
def test_stem_conv(U, cin, cout, H, W):
    g = torch.Generator().manual_seed(2)
    x = bf(torch.randn(2, cin, H, W, generator=g))
    w = torch.randn(cout, cin, 3, 3, generator=g) / 5
    ga, be = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    mu, va = torch.randn(cout, generator=g) * 0.1, torch.rand(cout, generator=g) + 0.5
    ws, bias = U.pack_stem(w.cuda(), (ga.cuda(), be.cuda(), mu.cuda(), va.cuda(), 1e-5))
    s = ga / torch.sqrt(va + 1e-5)
    ref = F.relu(F.conv2d(x, bf(w * s[:, None, None, None]), be - mu * s, padding=1))
    x4 = U.nchw_to_nhwc4(x.cuda())
    assert cin == 4 or x4[..., cin:].abs().max().item() == 0.0
    close(nchw(U.stem_conv(x4, ws, bias, cin)), bf(ref), 1e-2)


def test_head_logits_probs_mask(U):
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn(3, 64, 40, 56, generator=g))
    w, b = torch.randn(64, generator=g) / 8, 0.1
    lg, pr, mk = U.head(nhwc(x), w.cuda(), b, 0.5)
    ref = (x * w[None, :, None, None]).sum(1) + b
    close(lg, ref, 1e-5)
    close(pr, torch.sigmoid(ref), 1e-6)
    # the mask must be exactly what the reference post-process makes of the kernel's own probabilities
    want = O.postprocess_oracle([pr.cpu().numpy()[:1, None]], (40, 56), 0.5)
    assert np.array_equal(mk[0].cpu().numpy(), want)
    lg2, _, mk2 = U.head(nhwc(x), w.cuda(), b, 0.7, want=("logits", "mask"))
    assert torch.equal((mk2 > 0).cpu(), torch.sigmoid(lg2.cpu()) > 0.7) or \
        ((mk2 > 0).cpu() != (torch.sigmoid(lg2.cpu()) > 0.7)).float().mean() < 1e-5


@pytest.mark.parametrize("hs,ws", [(480, 640), (224, 224), (685, 1055), (960, 1280), (448, 448)])
def test_preprocess_is_cv2_exact(U, hs, ws):
    rng = np.random.default_rng(hs)
    img = rng.integers(0, 256, (2, hs, ws, 3), dtype=np.uint8)
    y, r = U.preprocess_u8(torch.from_numpy(img).cuda(), swap_rb=True, return_resized=True)
    want = np.stack([O.preprocess_oracle(im, (224, 224), swap_rb=True)[0][0] for im in img])
    assert np.array_equal(r.cpu().numpy(), want)                      # uint8 resize + BGR->RGB: bit-exact
    norm = torch.from_numpy(O.normalize_oracle(want)).permute(0, 2, 3, 1)
    assert (y[..., :3].float().cpu() - bf(norm)).abs().max().item() <= 2 ** -6   # one bf16 ulp at |x| < 4
    assert y[..., 3].abs().max().item() == 0.0


@pytest.mark.parametrize("hs,ws,h,w,b", [(37, 53, 64, 96, 3), (480, 640, 100, 60, 2), (1080, 1920, 224, 224, 1), (224, 224, 448, 448, 1),
                                         (200, 4001, 16, 16, 1), (17, 23, 8, 8, 5), (30, 42, 30, 42, 2), (32, 48, 32, 48, 3)])
def test_preprocess_tile_kernel_ragged_shapes(U, hs, ws, h, w, b):
    """The tile kernel's row staging (span / sparse modes, rows narrower than a 16-byte chunk, unaligned pitches, tiles that
    overhang the image, very wide rows that force fewer rows per CTA) stays bit-equal to cv2's resize."""
    rng = np.random.default_rng(hs * 7 + ws)
    img = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    _, r = U.preprocess_u8(torch.from_numpy(img).cuda(), size=(h, w), swap_rb=False, return_resized=True)
    want = np.stack([O.preprocess_oracle(im, (h, w))[0][0] for im in img])
    assert np.array_equal(r.cpu().numpy(), want)


def test_preprocess_golden_images(U, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for k in ("synthetic_480x640", "picture_684x1054", "frame_224x224"):
        _, r = U.preprocess_u8(torch.from_numpy(g[k + "_src"][None]).cuda(), return_resized=True)
        assert np.array_equal(r[0].cpu().numpy(), g[k + "_resized"]), k


def test_maxpool(U):
    x = bf(torch.randn(2, 64, 28, 36))
    assert torch.equal(nchw(U.maxpool2x2(nhwc(x))), F.max_pool2d(x, 2))


def _set(U, name, v):
    from unet_lane_detection_b200._lib import check, lib
    check(lib.unet_b200_set_option(name.encode(), v))


@pytest.mark.parametrize("hs,ws,pad,h,w,b", [(480, 640, 0, 224, 224, 3), (480, 640, 64, 224, 224, 2), (960, 1280, 0, 224, 224, 2),
                                             (1080, 1920, 16, 224, 224, 1), (100, 64, 0, 224, 224, 2), (131, 176, 32, 60, 300, 2),
                                             (300, 400, 0, 299, 400, 1), (9, 16, 0, 3, 5, 4), (480, 640, 0, 480, 640, 1),
                                             (685, 1055, 0, 224, 224, 2), (37, 53, 5, 64, 96, 3), (17, 23, 0, 8, 8, 5),
                                             (200, 4001, 3, 16, 16, 1)])
def test_preprocess_bulk_kernel_equals_cv2_and_thread_staged(U, hs, ws, pad, h, w, b):
    """preprocess_bulk_u8_kernel (rows staged by cp.async.bulk, two stages; 16-byte aligned frames take it by default):
    bit-equal to cv2 and to the thread-staged tile kernel, the normalised bf16 output included. Covers the span mode with
    ONE request per tile (contiguous rows), per-row requests (padded pitch), the sparse mode (down-scaling > 2x), the exact
    2x2 decimation, up-scaling, tiles that overhang the image, more output columns than threads, and rows that do not start
    on 16-byte boundaries (odd widths / pitches: aligned-down requests + per-slot offsets, tail bytes of the last frame copied
    by the requesting thread - the source tensor ends exactly at the last pixel)."""
    from unet_lane_detection_b200._lib import check, f3, lib
    from unet_lane_detection_b200.ops import MEAN_255, STD_255
    rng = np.random.default_rng(hs * 11 + ws + pad)
    img = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    pitch = ws * 3 + pad
    buf = torch.zeros(b, hs, pitch, dtype=torch.uint8)
    buf[:, :, :ws * 3] = torch.from_numpy(img.reshape(b, hs, ws * 3))
    buf = buf.cuda()
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for bulk in (1, 0):
        _set(U, "pre_bulk", bulk)
        try:
            y = torch.full((b, h, w, 4), 7.0, dtype=torch.bfloat16, device="cuda")
            r = torch.zeros(b, h, w, 3, dtype=torch.uint8, device="cuda")
            check(lib.unet_b200_preprocess_u8(buf.data_ptr(), b, hs, ws, pitch, hs * pitch, h, w, 0, f3(MEAN_255), f3(STD_255),
                                              y.data_ptr(), r.data_ptr(), st))
            torch.cuda.synchronize()
            outs.append((y.cpu(), r.cpu()))
        finally:
            _set(U, "pre_bulk", 1)
    want = np.stack([O.preprocess_oracle(im, (h, w))[0][0] for im in img])
    assert np.array_equal(outs[0][1].numpy(), want)
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])


@pytest.mark.parametrize("halo", [0, 1])
@pytest.mark.parametrize("B,H,W,C0,C1,Cout,pool", [(2, 224, 224, 64, 0, 64, True), (1, 224, 224, 64, 64, 64, False),
                                                   (2, 112, 112, 64, 0, 128, False), (2, 112, 112, 128, 0, 128, True),
                                                   (1, 112, 112, 128, 128, 128, False), (3, 24, 40, 64, 0, 64, True),
                                                   (1, 480, 640, 64, 0, 64, True), (1, 120, 160, 128, 0, 128, True)])
def test_conv3x3_halo_and_per_tap_kernels(U, halo, B, H, W, C0, C1, Cout, pool):
    """Both 3x3 kernels (halo patch + shifted descriptors / one TMA box per tap) on the shapes the plan routes to the
    halo kernel, including resident (<= 2 channel blocks) and streamed weights, partial tiles (H % 16 != 0)."""
    _set(U, "halo", halo)
    try:
        conv_case(U, B, H, W, C0, C1, Cout, pool=pool, seed=halo + 3)
    finally:
        _set(U, "halo", 1)


@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("cin,B,H,W", [(3, 2, 224, 224), (3, 1, 32, 48), (1, 3, 16, 16), (4, 1, 24, 40), (3, 1, 480, 640), (3, 5, 28, 36)])
def test_stem_conv_tensor_core(U, cin, B, H, W, wide):
    """Tensor-core stem (in-kernel im2col, K = 36 padded to 48) vs the oracle; partial tiles when H % 16 or W % 8 != 0.
    wide: the 4 x 32 tile form (one contiguous 4 KB row segment per TMA store) - same MMAs, bit-identical to the 16 x 8 form."""
    _set(U, "stem_wide", wide)
    try:
        _stem_case(U, cin, B, H, W)
    finally:
        _set(U, "stem_wide", 0)


def test_stem_tile_forms_are_bit_identical(U):
    x = U.nchw_to_nhwc4(bf(torch.randn(3, 3, 52, 72, generator=torch.Generator().manual_seed(8))).cuda())
    w = torch.randn(64, 3, 3, 3, generator=torch.Generator().manual_seed(9)) / 5
    wp, bias = U.pack_stem_tc(w.cuda(), None)
    outs = []
    for wide in (0, 1):
        _set(U, "stem_wide", wide)
        try:
            outs.append(U.stem_conv_tc(x, wp, bias))
        finally:
            _set(U, "stem_wide", 0)
    assert torch.equal(outs[0], outs[1])


def _stem_case(U, cin, B, H, W):
    g = torch.Generator().manual_seed(5)
    x = bf(torch.randn(B, cin, H, W, generator=g))
    w = torch.randn(64, cin, 3, 3, generator=g) / 5
    ga, be = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    mu, va = torch.randn(64, generator=g) * 0.1, torch.rand(64, generator=g) + 0.5
    wp, bias = U.pack_stem_tc(w.cuda(), (ga.cuda(), be.cuda(), mu.cuda(), va.cuda(), 1e-5))
    s = ga / torch.sqrt(va + 1e-5)
    ref = F.relu(F.conv2d(x, bf(w * s[:, None, None, None]), be - mu * s, padding=1))
    y = U.stem_conv_tc(U.nchw_to_nhwc4(x.cuda()), wp, bias)
    close(nchw(y), bf(ref), 1e-2)


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,pool", [(2, 224, 224, 64, 0, 64, True), (1, 224, 224, 64, 64, 64, False),
                                                   (2, 112, 112, 64, 0, 128, False), (2, 112, 112, 128, 0, 128, True),
                                                   (1, 112, 112, 128, 128, 128, False), (3, 24, 40, 64, 0, 64, True),
                                                   (1, 8, 24, 64, 0, 64, False), (5, 16, 8, 64, 0, 128, False),
                                                   (1, 120, 160, 128, 0, 128, True), (2, 56, 56, 256, 0, 64, False)])
def test_conv3x3_cta_pair_kernel_equals_single_cta(U, B, H, W, C0, C1, Cout, pool):
    """conv_halo2_kernel (cta_group::2: M = 256 over two SMs, half of the weight tile per SM) issues the same MMAs in the same
    order as conv_halo_kernel, so the outputs must be BIT-identical - on even and odd tile counts (the odd pair member is
    loaded out of bounds), resident and streamed weights, two K sources and the fused pool; and both match the oracle."""
    g = torch.Generator().manual_seed(11)
    x0 = nhwc(bf(torch.randn(B, C0, H, W, generator=g)))
    x1 = nhwc(bf(torch.randn(B, C1, H, W, generator=g))) if C1 else None
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=g) / (3.0 * (C0 + C1) ** 0.5)
    wp, bias = U.pack_conv3x3(w.cuda(), None)
    outs = []
    for pair in (1, 0):
        _set(U, "halo2", pair)
        try:
            o = U.conv3x3(x0, wp, bias, x1=x1, relu=True, pool=pool)
        finally:
            _set(U, "halo2", 1)
        torch.cuda.synchronize()
        outs.append(o if pool else (o,))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    conv_case(U, B, H, W, C0, C1, Cout, pool=pool, seed=5)


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,pool", [(2, 56, 56, 128, 0, 256, False), (2, 56, 56, 256, 0, 256, True),
                                                   (8, 28, 28, 256, 0, 512, True), (8, 28, 28, 512, 512, 512, False),
                                                   (32, 14, 14, 512, 0, 1024, False), (3, 14, 14, 64, 0, 256, False),
                                                   (5, 28, 28, 64, 64, 256, False), (1, 60, 80, 64, 0, 256, False),
                                                   (7, 8, 24, 128, 0, 512, False)])
def test_conv3x3_umma_cta_pair_equals_single_cta(U, B, H, W, C0, C1, Cout, pool):
    """conv_umma2_kernel (cta_group::2) against conv_umma_kernel on the per-tap implicit-GEMM layers: bit-identical outputs
    (odd and even pixel-tile counts, several column blocks, two K sources, fused pool, ragged sizes)."""
    g = torch.Generator().manual_seed(13)
    x0 = nhwc(bf(torch.randn(B, C0, H, W, generator=g)))
    x1 = nhwc(bf(torch.randn(B, C1, H, W, generator=g))) if C1 else None
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=g) / (3.0 * (C0 + C1) ** 0.5)
    wp, bias = U.pack_conv3x3(w.cuda(), None)
    outs = []
    for pair in (1, 0):
        _set(U, "umma2", pair)
        _set(U, "halo", 0)
        try:
            o = U.conv3x3(x0, wp, bias, x1=x1, relu=True, pool=pool)
        finally:
            _set(U, "umma2", 1)
            _set(U, "halo", 1)
        torch.cuda.synchronize()
        outs.append(o if pool else (o,))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    conv_case(U, B, H, W, C0, C1, Cout, pool=pool, seed=6)


@pytest.mark.parametrize("B,H,Cin,f", [(32, 14, 1024, 512), (2, 56, 256, 128), (1, 112, 128, 64), (3, 14, 128, 64)])
def test_convT2x2_cta_pair_equals_single_cta(U, B, H, Cin, f):
    g = torch.Generator().manual_seed(17)
    x = nhwc(bf(torch.randn(B, Cin, H, H, generator=g)))
    w = torch.randn(Cin, f, 2, 2, generator=g) / (Cin ** 0.5)
    b = torch.randn(f, generator=g).cuda()
    wp = U.pack_convT2x2(w.cuda())
    outs = []
    for pair in (1, 0):
        _set(U, "umma2", pair)
        try:
            outs.append(U.convT2x2(x, wp, b))
        finally:
            _set(U, "umma2", 1)
        torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
