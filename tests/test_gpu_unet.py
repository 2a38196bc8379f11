"""GPU (B200): the whole hot path through the drop-in module / executor against the oracle and the
golden fixtures generated from the reference listing."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
from conftest import record_parity

pytestmark = pytest.mark.gpu
LOGIT_TOL_BF16 = 2e-2   # BASELINE.json north_star: logits within 2e-2 abs on random-init weights (bf16 path)


@pytest.fixture(scope="module")
def U():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import unet_lane_detection_b200 as mod
    return mod


def make_pair(U, feats, gain=1.0, seed=0):
    torch.manual_seed(seed)
    ref = O.UNetOracle(3, 1, feats).eval()
    O.randomize_bn_(ref, seed=1)
    if gain != 1.0:
        O.scale_head_(ref, gain)
    net = U.UNet(3, 1, feats)
    net.load_state_dict(ref.state_dict())
    return ref, net.cuda().eval()


def test_golden_logits_from_reference_listing(U, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    for tag, feats in (("f64x2", [64, 128]), ("default", [64, 128, 256, 512])):
        _, net = make_pair(U, feats, gain=40.0)
        x = torch.from_numpy(g[f"{tag}_x"])
        with torch.no_grad():
            y = net(x.cuda()).cpu()
        ref = torch.from_numpy(g[f"{tag}_logits"])
        assert y.shape == ref.shape
        # head gain 40 makes |logit| ~ O(1): tolerance is relative to the logit range here
        assert (y - ref).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, ref.abs().max().item())


def test_full_size_random_init_parity(U):
    """configs[0] weights (seed-0 default init, BN randomised) at 224x224: the north_star gate."""
    ref, net = make_pair(U, [64, 128, 256, 512])
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        y32 = ref(x)
        y = net(x.cuda()).cpu()
    assert y.shape == (2, 1, 224, 224) and y.dtype == torch.float32
    assert (y - y32).abs().max().item() <= LOGIT_TOL_BF16
    # mask agreement: random-init logits hug zero, so report against the bf16 noise floor of the oracle itself
    yem = O.forward_bf16_emulated(ref, x)
    floor = O.mask_agreement(yem, y32)
    got = O.mask_agreement(y, y32)
    assert got >= min(0.999, floor - 0.002), f"mask agreement {got:.5f} (bf16-emulated oracle floor {floor:.5f})"
    band = 4 * (yem - y32).abs().max().item()
    banded = O.mask_agreement(y, y32, band=band)
    assert banded >= 0.9999
    # the noise-floor row SURVEY.md 7 asks for, recorded (profiles/r2_parity.json)
    record_parity("random_init_gain1_batch2_224", {
        "max_abs_logit_err": (y - y32).abs().max().item(), "logit_abs_max": y32.abs().max().item(), "logit_std": y32.std().item(),
        "mask_agreement": got, "mask_agreement_bf16_emulated_oracle": floor, "emulated_max_abs_logit_err": (yem - y32).abs().max().item(),
        "mask_agreement_outside_band": banded, "band": band,
        "mask_agreement_excluding_abs_logit_below_1e-3": O.mask_agreement(y, y32, band=1e-3),
        "gate": "logits <= 2e-2 abs; agreement >= min(0.999, emulated floor - 0.002); >= 0.9999 outside 4x the emulated max error"})


def test_full_size_realistic_logit_scale(U):
    """Same weights with the head scaled so |logit| ~ O(1): >= 99.9 % thresholded-mask agreement (north_star)."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(99))
    with torch.no_grad():
        y32 = ref(x)
        y = net(x.cuda()).cpu()
    assert (y - y32).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, y32.abs().max().item())
    got = O.mask_agreement(y, y32)
    assert got >= 0.999
    yem = O.forward_bf16_emulated(ref, x)
    record_parity("random_init_gain40_batch2_224", {
        "max_abs_logit_err": (y - y32).abs().max().item(), "logit_abs_max": y32.abs().max().item(), "logit_std": y32.std().item(),
        "mask_agreement": got, "mask_agreement_bf16_emulated_oracle": O.mask_agreement(yem, y32),
        "emulated_max_abs_logit_err": (yem - y32).abs().max().item(),
        "gate": "logits <= 2e-2 * max(1, |z|max) (the 2e-2 abs bound of north_star is stated for |z| ~ 1e-2 random-init logits; "
                "with the head scaled x40 the same relative error is 2.9e-2 abs at |z|max 4); agreement >= 0.999"})


def test_whole_net_batch64_camera_frames_against_oracle(U):
    """the batch-folded tiles the benchmark runs (TB = 8 at 28x28, TB = 32 at 14x14 are only live from batch 8 / 32
    up) end to end against the oracle pipeline, on 480x640 camera frames (a real resize, src/unet.py:33), one 64-frame chunk."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    net.b200_chunk = 64
    rng = np.random.default_rng(64)
    frames = rng.integers(0, 256, (64, 480, 640, 3), dtype=np.uint8)
    logits, _, mask = net.predict_mask(torch.from_numpy(frames).cuda(), threshold=0.5, swap_rb=True, want=("logits", "mask"))
    pre = np.concatenate([O.preprocess_oracle(f, (224, 224), swap_rb=True)[0] for f in frames])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        z = torch.cat([ref(torch.from_numpy(O.normalize_oracle(pre[i:i + 16]))) for i in range(0, 64, 16)])
    err = (logits.cpu() - z[:, 0]).abs()
    rngz = max(1.0, z.abs().max().item())
    assert err.max().item() <= LOGIT_TOL_BF16 * rngz, err.max().item()
    want = np.stack([O.postprocess_oracle([z[i:i + 1].numpy()], (224, 224), 0.5) for i in range(64)])
    agree = float((mask.cpu().numpy() == want).mean())
    assert agree >= 0.999, agree
    per_frame = err.reshape(64, -1).max(dim=1).values
    assert per_frame.max().item() <= 3.0 * per_frame.median().item() + 1e-3      # no frame position of the folded tiles stands out
    record_parity("batch64_chunk64_480x640_sources_gain40", {
        "max_abs_logit_err": err.max().item(), "logit_abs_max": z.abs().max().item(), "mask_agreement": agree,
        "per_frame_max_err_median": per_frame.median().item(), "per_frame_max_err_max": per_frame.max().item()})
    # the same frames in four chunks of 16 (different tile folding at the deep levels) give the same logits
    net.b200_chunk = 16
    l16, _, m16 = net.predict_mask(torch.from_numpy(frames).cuda(), threshold=0.5, swap_rb=True, want=("logits", "mask"))
    assert torch.equal(l16, logits) and torch.equal(m16, mask)


def test_out_channels_greater_than_one(U):
    """README.md:1447 builds nn.Conv2d(features[0], out_channels, 1) for any out_channels: the plan accepts them (the head runs
    as its own kernel) and the module returns [B, out_channels, H, W] like the reference."""
    torch.manual_seed(3)
    ref = O.UNetOracle(3, 3, [64, 128]).eval()
    O.randomize_bn_(ref, seed=1)
    with torch.no_grad():
        ref.output.weight.mul_(40.0)
    net = U.UNet(3, 3, [64, 128])
    net.load_state_dict(ref.state_dict())
    net = net.cuda().eval()
    x = torch.randn(5, 3, 32, 48, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = ref(x)
        got = net(x.cuda()).cpu()
    assert got.shape == (5, 3, 32, 48)
    assert (got - want).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, want.abs().max().item())
    frames = torch.randint(0, 256, (5, 32, 48, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(6))
    logits, probs, mask = net.predict_mask(frames.cuda(), size=(32, 48), want=("logits", "probs", "mask"))
    assert logits.shape == probs.shape == mask.shape == (5, 3, 32, 48)
    assert torch.equal(mask, (probs > 0.5).to(torch.uint8) * 255)
    assert (probs - torch.sigmoid(logits)).abs().max().item() < 1e-6
    # host-buffer entry with three output planes per frame
    m_host = torch.empty(5, 3, 32, 48, dtype=torch.uint8).pin_memory()
    net.infer_host(frames.pin_memory(), size=(32, 48), mask_out=m_host)
    assert torch.equal(m_host, mask.cpu())


def test_pipeline_matches_reference_pipeline(U):
    """uint8 camera frames -> preprocess -> net -> sigmoid -> >thr*255, vs src/unet.py:24-72 restated in the oracle."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (3, 480, 640, 3), dtype=np.uint8)
    logits, probs, mask = net.predict_mask(torch.from_numpy(frames).cuda(), threshold=0.5, swap_rb=True,
                                           want=("logits", "probs", "mask"))
    pre = np.concatenate([O.preprocess_oracle(f, (224, 224), swap_rb=True)[0] for f in frames])
    with torch.no_grad():
        z = ref(torch.from_numpy(O.normalize_oracle(pre)))
    assert (logits.cpu() - z[:, 0]).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, z.abs().max().item())
    want = np.stack([O.postprocess_oracle([z[i:i + 1].numpy()], (224, 224), 0.5) for i in range(3)])
    agree = (mask.cpu().numpy() == want).mean()
    assert agree >= 0.999, agree
    # the mask is bit-exactly the reference post-process applied to the kernel's own probabilities
    own = np.stack([O.postprocess_oracle([probs[i:i + 1, None].cpu().numpy()], (224, 224), 0.5) for i in range(3)])
    assert np.array_equal(mask.cpu().numpy(), own)
    assert set(np.unique(mask.cpu().numpy())) <= {0, 255}


def test_chunking_and_permutation_invariance(U):
    """Size-independent properties at a larger batch: results do not depend on how the batch is chunked or ordered."""
    _, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    frames = torch.randint(0, 256, (21, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(7)).cuda()
    net.b200_chunk = 32
    l_a, _, m_a = net.predict_mask(frames, want=("logits", "mask"))
    net.b200_chunk = 8                      # 21 = 8 + 8 + 5: ragged last chunk
    l_b, _, m_b = net.predict_mask(frames, want=("logits", "mask"))
    assert torch.equal(l_a, l_b) and torch.equal(m_a, m_b)
    perm = torch.randperm(21, generator=torch.Generator().manual_seed(8)).cuda()
    l_c, _, _ = net.predict_mask(frames[perm].contiguous(), want=("logits",))
    assert torch.equal(l_c, l_b[perm])
    net.b200_chunk = 32


def test_batch_one_and_weight_updates(U):
    ref, net = make_pair(U, [64, 128], gain=40.0)
    x = torch.randn(1, 3, 32, 32)
    with torch.no_grad():
        y0 = net(x.cuda()).cpu()
        assert (y0 - ref(x)).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, ref(x).abs().max().item())
        # in-place parameter updates (optimizer.step / load_state_dict) must invalidate the packed weights
        ref.output.bias.add_(1.0)
        net.load_state_dict(ref.state_dict())
        y1 = net(x.cuda()).cpu()
    assert (y1 - y0 - 1.0).abs().max().item() < 1e-5


def test_camera_resolution_config5_smoke(U):
    """configs[4] geometry (480x640) on a small batch: same kernels, different tensor maps."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    x = torch.randn(1, 3, 480, 640, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        y32 = ref(x)
        y = net(x.cuda()).cpu()
    assert (y - y32).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, y32.abs().max().item())


def test_executor_contract(U, tmp_path):
    """RKNN_model_container contract (src/py_utils/rknn_executor.py:26-42): list in, list of ndarray out, release()."""
    ref, _ = make_pair(U, [64, 128, 256, 512], gain=40.0)
    path = tmp_path / "best_model.pth"
    torch.save({"epoch": 1, "model_state_dict": ref.state_dict()}, path)
    box = U.B200_model_container(str(path))
    rng = np.random.default_rng(3)
    frame = rng.integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
    out = box.run([frame])
    assert isinstance(out, list) and out[0].shape == (1, 1, 224, 224) and out[0].dtype == np.float32
    assert out[0].min() >= 0.0 and out[0].max() <= 1.0                 # probabilities: sigmoid is inside the graph
    assert np.array_equal(box.run(frame)[0], out[0])                    # bare array is wrapped (rknn_executor.py:31-34)
    with torch.no_grad():
        z = ref(torch.from_numpy(O.normalize_oracle(frame)))
    mask_ref = O.postprocess_oracle([torch.sigmoid(z).numpy()], (224, 224), 0.5)
    mask_got = O.postprocess_oracle(out, (224, 224), 0.5)               # the caller's own post-process on our output
    assert (mask_ref == mask_got).mean() >= 0.999
    box.release()
    assert box.run([frame]) == []                                       # run after release (rknn_executor.py:27-29)


def test_infer_host_equals_device_path(U):
    _, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    frames = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    _, _, m_dev = net.predict_mask(frames.cuda(), want=("mask",))
    m_host = torch.empty(5, 224, 224, dtype=torch.uint8).pin_memory()
    net.infer_host(frames.pin_memory(), mask_out=m_host)
    assert torch.equal(m_host, m_dev.cpu())


@pytest.mark.parametrize("hs,ws", [(120, 160), (480, 640)])
def test_infer_host_pieces_and_passes(U, hs, ws):
    """unet_b200_infer_u8_host_stream: a pass is cut into pieces whose copies / preprocess / first and last layers are
    pipelined (include/unet_b200.h). 77 camera frames through a plan of 32 (passes 32 / 32 / 13, two pieces each, ragged last
    piece): logits and masks are bit-equal to the device-resident path, for the piece-wise and the pass-granular schedule.
    480x640 frames are "copy-bound" (more than 1.5x the network input): they take the hybrid schedule (short first pass of 8,
    then 32 / 32 / 5, pieces inside) or, with host_hybrid = 0, the pass-granular one. Pieces are geometric by default (16, 16
    for a pass of 32; 16, 32, 29 for the 77-frame pass of a plan of 80, the last layer running them largest first) or, with
    host_geometric = 0, equal."""
    from unet_lane_detection_b200._lib import check, lib
    ref, _ = make_pair(U, [64, 128, 256, 512], gain=40.0)
    frames = torch.randint(0, 256, (77, hs, ws, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(6))
    outs = {}
    try:
        for pieces, hybrid, geometric, chunk in ((8, 1, 1, 32), (8, 1, 1, 80), (0, 1, 1, 32), (8, 0, 1, 32), (8, 1, 0, 32)):
            check(lib.unet_b200_set_option(b"host_pieces", pieces))
            check(lib.unet_b200_set_option(b"host_hybrid", hybrid))
            check(lib.unet_b200_set_option(b"host_geometric", geometric))
            net = U.UNet(3, 1, [64, 128, 256, 512])
            net.load_state_dict(ref.state_dict())
            net = net.cuda().eval()
            net.b200_chunk = chunk
            m_host = torch.zeros(77, 224, 224, dtype=torch.uint8).pin_memory()
            l_host = torch.zeros(77, 224, 224, dtype=torch.float32).pin_memory()
            net.infer_host(frames.pin_memory(), swap_rb=True, mask_out=m_host, logits_out=l_host)
            net.infer_host(frames.pin_memory(), swap_rb=True, mask_out=m_host, logits_out=l_host)   # events / slots are reused
            outs[(pieces, hybrid, geometric, chunk)] = (m_host.clone(), l_host.clone(), net.gpu_launches)
            if pieces and hybrid and geometric and chunk == 32:
                l_dev, _, m_dev = net.predict_mask(frames.cuda(), swap_rb=True, want=("logits", "mask"))
    finally:
        check(lib.unet_b200_set_option(b"host_pieces", 8))
        check(lib.unet_b200_set_option(b"host_hybrid", 1))
        check(lib.unet_b200_set_option(b"host_geometric", 1))
    for key, (m, l, _) in outs.items():
        assert torch.equal(m, m_dev.cpu()), key
        assert torch.equal(l, l_dev.cpu()), key


@pytest.mark.parametrize("feats,B,H,W", [([64, 128, 256, 512], 3, 224, 224), ([64, 128], 5, 40, 56), ([32, 64], 1, 16, 8),
                                         ([64, 128], 2, 72, 200)])
def test_stem_fused_into_second_conv_is_bit_identical(U, feats, B, H, W):
    """Option stem_fuse (off by default - same speed): the stem conv runs inside the patch producer of enc0.conv1 (stem_halo2_kernel) and its
    output never goes to HBM. Same MMAs and rounding points as the two-kernel path: logits are bit-equal, including ragged
    tiles (rows 40 = 2 x 16 + 8), an odd tile count (the CTA pair's missing tile), a single pair and zero-extended widths."""
    from unet_lane_detection_b200._lib import check, lib
    ref, _ = make_pair(U, feats, gain=40.0)
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(8)).cuda()
    outs, launches = [], []
    try:
        for fuse in (1, 0):
            check(lib.unet_b200_set_option(b"stem_fuse", fuse))
            net = U.UNet(3, 1, feats)
            net.load_state_dict(ref.state_dict())
            net = net.cuda().eval()
            with torch.no_grad():
                outs.append(net(x).clone())
            launches.append(net.gpu_launches)
    finally:
        check(lib.unet_b200_set_option(b"stem_fuse", 0))
    assert torch.equal(outs[0], outs[1])
    assert launches[0] == launches[1] - 1          # one kernel fewer
    with torch.no_grad():
        want = ref(x.cpu())
    assert (outs[0].cpu() - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())


def test_fused_head_and_halo_switches_agree(U):
    """The fused (1x1 head in the last conv's epilogue) and unfused paths, and the two 3x3 kernels, give the same net."""
    from unet_lane_detection_b200._lib import check, lib
    ref, _ = make_pair(U, [64, 128, 256, 512], gain=40.0)
    frames = torch.randint(0, 256, (3, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(4)).cuda()
    outs = {}
    try:
        for halo, fuse in ((1, 1), (1, 0), (0, 0)):
            check(lib.unet_b200_set_option(b"halo", halo))
            check(lib.unet_b200_set_option(b"fuse_head", fuse))
            net = U.UNet(3, 1, [64, 128, 256, 512])
            net.load_state_dict(ref.state_dict())
            net = net.cuda().eval()
            outs[(halo, fuse)] = net.predict_mask(frames, want=("logits", "probs", "mask"))
    finally:
        check(lib.unet_b200_set_option(b"halo", 1))
        check(lib.unet_b200_set_option(b"fuse_head", 1))
    la, pa, ma = outs[(1, 1)]
    lb, pb, mb = outs[(1, 0)]
    lc, _, mc = outs[(0, 0)]
    assert (la - lb).abs().max().item() < 1e-4          # same activations, different summation order in the head
    assert (ma != mb).float().mean().item() < 1e-4
    assert (la - lc).abs().max().item() < 2e-2 * max(1.0, lc.abs().max().item())   # different K order in the convs
    assert (ma != mc).float().mean().item() < 2e-3
    own = (pa > 0.5).to(torch.uint8) * 255
    assert torch.equal(ma, own)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_small_plan_column_blocks_are_bit_identical(U, precision):
    """Option small_n (default on): a plan too small to give every SM a 128 x 256 tile - the per-frame executor path - runs
    its per-tap layers with BLOCK_N 128 / 64 (more CTAs). Every output element still accumulates the same products in the
    same order, so the logits are bit-equal to the BLOCK_N = 256 plan (which the small-batch tests no longer exercise
    end to end otherwise), for the bf16 and the fp32-class plan."""
    from unet_lane_detection_b200._lib import check, lib
    ref, _ = make_pair(U, [64, 128, 256, 512], gain=40.0)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(11)).cuda()
    outs = []
    try:
        for small in (1, 0):
            check(lib.unet_b200_set_option(b"small_n", small))
            net = U.UNet(3, 1, [64, 128, 256, 512])
            net.load_state_dict(ref.state_dict())
            net = net.cuda().eval()
            net.b200_precision = precision
            with torch.no_grad():
                outs.append(net(x).clone())
    finally:
        check(lib.unet_b200_set_option(b"small_n", 1))
    assert torch.equal(outs[0], outs[1])
    with torch.no_grad():
        want = ref(x.cpu())
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert (outs[0].cpu() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("feats,H,W", [([48, 96], 64, 96), ([24, 48, 96], 64, 64), ([96, 192], 32, 64), ([20], 16, 24),
                                       ([72, 144, 288], 32, 32)])
def test_arbitrary_feature_widths(U, feats, H, W):
    """UNet(features=...) takes any list of widths the reference's own module accepts (README.md:1424; its decoder needs every
    width to be twice the one before): channel counts that are not multiples of 64 are stored
    zero-extended (zero weights, zero bias, relu(0) = 0), so the logits are those of the logical network. Covers widths
    below and above 64 in the first block (tensor-core stem / FP32-pipe stem), odd multiples of 8 and a width above 256."""
    ref, net = make_pair(U, feats, gain=40.0)
    x = torch.randn(3, 3, H, W, generator=torch.Generator().manual_seed(sum(feats)))
    with torch.no_grad():
        got = net(x.cuda()).cpu()
        want = ref(x)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    frames = torch.randint(0, 256, (2, H, W, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    _, _, mask = net.predict_mask(frames.cuda(), size=(H, W), want=("mask",))
    with torch.no_grad():
        z = ref(torch.from_numpy(O.normalize_oracle(frames.numpy())))
    agree = ((z[:, 0] > 0).numpy() == (mask.cpu().numpy() > 0)).mean()
    assert agree >= 0.995, agree


def test_deployed_topology_32_64_128(U, golden_dir, tmp_path):
    """the topology of model/lane_unet*.rknn = UNet(features=[32,64,128]), 1,927,009
    parameters. Widths that are not multiples of 64 are stored zero-extended; logits must match the oracle as for the
    default network, through the module and through the executor with a milesial-named checkpoint."""
    ref, net = make_pair(U, [32, 64, 128], gain=40.0)
    assert sum(p.numel() for p in net.parameters()) == 1927009
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        y32 = ref(x)
        y = net(x.cuda()).cpu()
    assert (y - y32).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, y32.abs().max().item())
    # the narrow network's logits hug the threshold more than the default one's: gate against the bf16 noise floor of the
    # oracle itself (PyTorch with bf16 rounding at the same points), as for the random-init test above
    yem = O.forward_bf16_emulated(ref, x)
    floor = O.mask_agreement(yem, y32)
    got = O.mask_agreement(y, y32)
    assert got >= min(0.999, floor - 0.002), f"mask agreement {got:.5f} (bf16-emulated oracle floor {floor:.5f})"
    assert O.mask_agreement(y, y32, band=4 * (yem - y32).abs().max().item()) >= 0.9999
    # unfused head / non-halo kernels take the padded channels too
    from unet_lane_detection_b200._lib import check, lib
    try:
        check(lib.unet_b200_set_option(b"halo", 0))
        check(lib.unet_b200_set_option(b"fuse_head", 0))
        net2 = U.UNet(3, 1, [32, 64, 128])
        net2.load_state_dict(ref.state_dict())
        with torch.no_grad():
            y2 = net2.cuda().eval()(x.cuda()).cpu()
    finally:
        check(lib.unet_b200_set_option(b"halo", 1))
        check(lib.unet_b200_set_option(b"fuse_head", 1))
    assert (y2 - y32).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, y32.abs().max().item())
    # golden logits of the reference listing for this topology are pinned by the manifest's parameter count; executor path:
    sd = ref.state_dict()
    names = {"encoder_blocks.0": "inc", "encoder_blocks.1": "down1", "encoder_blocks.2": "down2", "bottleneck": "down3",
             "decoder_blocks.0": "up1", "decoder_blocks.1": "conv1", "decoder_blocks.2": "up2", "decoder_blocks.3": "conv2",
             "decoder_blocks.4": "up3", "decoder_blocks.5": "conv3", "output": "outc"}
    renamed = {}
    for k, v in sd.items():
        pre = next(p for p in sorted(names, key=len, reverse=True) if k.startswith(p + "."))
        renamed[names[pre] + k[len(pre):]] = v
    path = tmp_path / "lane_unet_deployed.pth"
    torch.save(renamed, path)
    box = U.B200_model_container(str(path))
    frame = np.random.default_rng(4).integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
    out = box.run([frame])[0]
    with torch.no_grad():
        want = torch.sigmoid(ref(torch.from_numpy(O.normalize_oracle(frame)))).numpy()
    assert out.shape == (1, 1, 224, 224) and np.abs(out - want).max() <= 2e-2
    box.release()


# ---------------------------------------------------------------- fp32-class path (split-bf16 x3 on the tensor cores)
LOGIT_TOL_FP32 = 1e-4   # BASELINE.json north_star: logits within 1e-4 on the fp32 path


def _split(t):
    """fp32 NHWC -> bf16 [.., 2C] = [hi | lo]."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous()


def _join(t):
    c = t.shape[-1] // 2
    return t[..., :c].float() + t[..., c:].float()


def test_split_layers_against_torch_fp32(U):
    """Every split-precision layer entry point against the fp32 torch op on the values the kernel actually sees."""
    import torch.nn.functional as F
    ops = U.ops
    g = torch.Generator().manual_seed(7)
    B, H, W = 3, 12, 20
    x0 = torch.randn(B, H, W, 64, generator=g)
    x1 = torch.randn(B, H, W, 128, generator=g)
    w = torch.randn(128, 192, 3, 3, generator=g) * 0.05
    s0, s1 = _split(x0).cuda(), _split(x1).cuda()
    wp, bias = ops.pack_conv3x3_split(w.cuda(), None, c0=64)
    y = _join(ops.conv3x3_split(s0, wp, bias, x1=s1, relu=True).cpu())
    xin = torch.cat([_join(s0.cpu()), _join(s1.cpu())], dim=-1).permute(0, 3, 1, 2).double()
    ref = F.relu(F.conv2d(xin, w.double(), padding=1)).permute(0, 2, 3, 1).float()
    assert (y - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    # ConvT
    wt = torch.randn(128, 64, 2, 2, generator=g) * 0.1
    bt = torch.randn(64, generator=g)
    up = _join(ops.convT2x2_split(s1, ops.pack_convT2x2_split(wt.cuda()), bt.cuda()).cpu())
    ref = F.conv_transpose2d(_join(s1.cpu()).permute(0, 3, 1, 2).double(), wt.double(), bt.double(), stride=2).permute(0, 2, 3, 1).float()
    assert up.shape == ref.shape
    assert (up - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    # pool: exact on the joined values
    pl = _join(ops.maxpool2x2_split(s0).cpu())
    ref = F.max_pool2d(_join(s0.cpu()).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(pl, ref)
    # stem from the fp32 NCHW input
    xs = torch.randn(B, 3, H, W, generator=g)
    wsx = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    ws, sb = ops.pack_stem(wsx.cuda(), None, fp32=True)
    st = _join(ops.stem_conv_split(xs.cuda(), ws, sb, relu=True).cpu())
    ref = F.relu(F.conv2d(xs.double(), wsx.double(), padding=1)).permute(0, 2, 3, 1).float()
    assert (st - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("gain", [1.0, 40.0])
def test_fp32_path_logits_within_1e4(U, gain):
    """b200_precision = 'fp32': logits within 1e-4 of the fp32 reference forward (north_star gate of the fp32 path);
    the fp64 forward of the same weights is the yardstick that shows how much of that is the reference's own rounding."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=gain)
    net.b200_precision = "fp32"
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(4321))
    with torch.no_grad():
        y32 = ref(x)
        y64 = ref.double()(x.double()).float()
        ref.float()
        y = net(x.cuda()).cpu()
    scale = max(1.0, y32.abs().max().item())
    err = (y - y32).abs().max().item()
    print(f"fp32 path: max|err| vs fp32 oracle {err:.3e}, vs fp64 {(y - y64).abs().max().item():.3e}; "
          f"fp32 oracle vs fp64 {(y32 - y64).abs().max().item():.3e}; |logit| max {scale:.2f}")
    assert err <= LOGIT_TOL_FP32 * scale
    assert O.mask_agreement(y, y32) >= 0.999
    net.b200_precision = "bf16"
    with torch.no_grad():
        yb = net(x.cuda()).cpu()
    assert (yb - y32).abs().max().item() > err   # the bf16 path is the coarser one (sanity: the switch does something)


def _torch_model_container_run(path, input_datas):
    """The reference's Torch_model_container (src/py_utils/pytorch_executor.py:14-61), its load and run() steps in order:
    torch.jit.load -> eval -> torch.tensor(ndarray) (CPU tensors) -> model(*inputs) -> dequantize -> .cpu().detach().numpy()."""
    pt_model = torch.jit.load(path)
    pt_model.eval()
    ins = [torch.tensor(d) for d in input_datas]
    ins = [v.float() if v.dtype == torch.float64 else v for v in ins]
    result = pt_model(*ins)
    result = list(result) if isinstance(result, tuple) else result
    result = result if isinstance(result, list) else [result]
    return [torch.dequantize(r).cpu().detach().numpy() for r in result]


def test_torchscript_trace_and_torch_model_container(U, tmp_path):
    """torch.jit.trace(model) serialises (the graph holds one unet_b200::infer operator and the model
    state as a constant) and the reference's Torch_model_container flow runs the file - float NCHW module and the uint8 NHWC
    lane graph (normalisation + sigmoid inside, like the deployed RKNN blob)."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(31))
    with torch.no_grad():
        eager = net(x.cuda()).cpu()
    traced = torch.jit.trace(net, x.cuda(), check_trace=False)
    assert "unet_b200::infer" in str(traced.graph)
    assert torch.equal(traced(x.cuda()).cpu(), eager)
    p1 = str(tmp_path / "unet_float.pt")
    torch.jit.save(traced, p1)
    out = _torch_model_container_run(p1, [x.numpy()])                  # CPU ndarray in, as the container feeds it
    assert len(out) == 1 and out[0].shape == (2, 1, 224, 224)
    assert np.array_equal(out[0], eager.numpy())
    other = _torch_model_container_run(p1, [x[:1, :, :64, :96].numpy().copy()])   # a trace is not tied to the example's shape
    with torch.no_grad():
        assert np.array_equal(other[0], net(x[:1, :, :64, :96].contiguous().cuda()).cpu().numpy())
    # the lane node's contract: uint8 NHWC RGB (1,224,224,3) -> probabilities (1,1,224,224)  (src/unet.py:24-72)
    p2 = str(tmp_path / "lane_unet_b200.pt")
    U.export_torchscript(net, p2, input_kind="uint8_nhwc", sigmoid=True)
    frame = np.random.default_rng(2).integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
    probs = _torch_model_container_run(p2, [frame])[0]
    with torch.no_grad():
        want = torch.sigmoid(ref(torch.from_numpy(O.normalize_oracle(frame)))).numpy()
    assert probs.shape == (1, 1, 224, 224) and probs.dtype == np.float32
    assert np.abs(probs - want).max() <= 2e-2
    mask = O.postprocess_oracle([probs], (224, 224), 0.5)               # the reference post-process on the container's output
    assert (mask == O.postprocess_oracle([want], (224, 224), 0.5)).mean() >= 0.999


def test_fp32_plan_from_uint8_frames_and_executor(U, tmp_path):
    """The fp32-class path is a plan of its own (UB_PRECISION_FP32): reachable from predict_mask / infer_host (fp32 preprocess,
    the image is never rounded to bf16) and from the executor, where the per-frame call replays a captured CUDA graph."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    net.b200_precision = "fp32"
    rng = np.random.default_rng(17)
    frames = rng.integers(0, 256, (3, 480, 640, 3), dtype=np.uint8)
    logits, probs, mask = net.predict_mask(torch.from_numpy(frames).cuda(), swap_rb=True, want=("logits", "probs", "mask"))
    pre = np.concatenate([O.preprocess_oracle(f, (224, 224), swap_rb=True)[0] for f in frames])
    with torch.no_grad():
        z = ref(torch.from_numpy(O.normalize_oracle(pre)))
    err = (logits.cpu() - z[:, 0]).abs().max().item()
    assert err <= LOGIT_TOL_FP32 * max(1.0, z.abs().max().item()), err
    want = np.stack([O.postprocess_oracle([z[i:i + 1].numpy()], (224, 224), 0.5) for i in range(3)])
    assert (mask.cpu().numpy() == want).mean() >= 0.9999
    # host-buffer entry point picks the fp32 preprocess from the plan's precision
    m_host = torch.empty(3, 224, 224, dtype=torch.uint8).pin_memory()
    l_host = torch.empty(3, 224, 224, dtype=torch.float32).pin_memory()
    net.infer_host(torch.from_numpy(frames).pin_memory(), swap_rb=True, mask_out=m_host, logits_out=l_host)
    assert torch.equal(m_host, mask.cpu()) and torch.equal(l_host, logits.cpu())
    record_parity("fp32_plan_480x640_sources_gain40", {"max_abs_logit_err": err, "logit_abs_max": z.abs().max().item(),
                                                       "gate": "<= 1e-4 * max(1, |z|max)"})
    # executor: graph replay of the fp32 plan (same frame twice -> identical output), probabilities within 1e-4
    path = tmp_path / "m.pth"
    torch.save(ref.state_dict(), path)
    box = U.B200_model_container(str(path), precision="fp32")
    rgb = np.ascontiguousarray(pre[:1])
    a = box.run([rgb])[0]
    b = box.run([rgb])[0]
    assert np.array_equal(a, b)
    with torch.no_grad():
        want_p = torch.sigmoid(ref(torch.from_numpy(O.normalize_oracle(rgb)))).numpy()
    assert np.abs(a - want_p).max() <= 1e-4
    box.release()
    net.b200_precision = "bf16"


@pytest.mark.parametrize("feats,B,H,W", [([64, 128], 3, 64, 96), ([64, 128, 256], 1, 40, 56)])
def test_fp32_path_other_topologies(U, feats, B, H, W):
    """fp32-class path on shallower networks, non-square inputs and odd batches (partial tiles, odd pixel-tile counts)."""
    ref, net = make_pair(U, feats, gain=20.0)
    net.b200_precision = "fp32"
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        y32 = ref(x)
        y = net(x.cuda()).cpu()
    assert y.shape == y32.shape
    assert (y - y32).abs().max().item() <= LOGIT_TOL_FP32 * max(1.0, y32.abs().max().item())


def test_camera_resolution_parity(U):
    """configs[4] geometry (480x640, no resize) at batch 1: the small-batch image-sized tiles (pick_tile) on every level,
    including the 30x40 bottleneck whose tiles do not divide the image."""
    ref, net = make_pair(U, [64, 128, 256, 512], gain=40.0)
    x = torch.randn(1, 3, 480, 640, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        y32 = ref(x)
        y = net(x.cuda()).cpu()
    assert (y - y32).abs().max().item() <= LOGIT_TOL_BF16 * max(1.0, y32.abs().max().item())
    assert O.mask_agreement(y, y32) >= 0.999
