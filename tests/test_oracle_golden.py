"""CPU: the oracle (oracle/unet_oracle.py) against the fixtures generated from the reference itself
(tests/golden/make_golden.py executes the reference's README listing and src/unet.py bodies)."""
import json
import os

import numpy as np
import torch

from oracle import unet_oracle as O


def _model(feats, gain=40.0):
    torch.manual_seed(0)
    m = O.UNetOracle(3, 1, feats).eval()
    O.randomize_bn_(m, seed=1)
    O.scale_head_(m, gain)
    return m


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


def test_manifest_matches_reference_listing(golden_dir):
    man = json.load(open(os.path.join(golden_dir, "unet_manifest.json")))
    for name in ("default", "deployed"):
        ent = man[name]
        torch.manual_seed(0)
        m = O.UNetOracle(3, 1, ent["features"])
        sd = m.state_dict()
        assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == ent["entries"]
        assert sum(p.numel() for p in m.parameters()) == ent["n_params"]
        assert len(sd) == ent["n_entries"]
        assert abs(_checksum(sd) - ent["abs_checksum_seed0"]) < 1e-6 * ent["abs_checksum_seed0"]
    assert man["default"]["n_params"] == 31037633      # reference README.md:2288
    assert man["default"]["n_entries"] == 118
    assert man["deployed"]["n_params"] == 1927009      # SURVEY.md D3 (topology of the shipped .rknn blobs)


def test_logits_match_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    for tag, feats in (("f64x2", [64, 128]), ("default", [64, 128, 256, 512])):
        m = _model(feats)
        assert abs(_checksum(m.state_dict()) - float(g[f"{tag}_checksum"])) < 1e-9 * float(g[f"{tag}_checksum"]), \
            "seeded initialiser drifted: regenerate tests/golden with make_golden.py"
        with torch.no_grad():
            y = m(torch.from_numpy(g[f"{tag}_x"]))
        np.testing.assert_allclose(y.numpy(), g[f"{tag}_logits"], rtol=0, atol=1e-5)


def test_output_shape_contract():
    m = O.UNetOracle(3, 1, [64, 128]).eval()    # reference README.md:1487-1490 (shape contract), small widths for speed
    with torch.no_grad():
        assert m(torch.randn(1, 3, 32, 32)).shape == (1, 1, 32, 32)


def test_loss_matches_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_small.npz"))
    logits = torch.from_numpy(g["logits"]).requires_grad_(True)
    crit = O.BCEDiceLossOracle(0.5, 0.5, pos_weight=torch.tensor([3.0]))
    total, bce, dice = crit(logits, torch.from_numpy(g["target"]))
    total.backward()
    assert abs(total.item() - float(g["total"])) < 1e-6
    assert abs(bce.item() - float(g["bce"])) < 1e-6
    assert abs(dice.item() - float(g["dice"])) < 1e-6
    np.testing.assert_allclose(logits.grad.numpy(), g["grad"], atol=1e-8)


def test_resize_matches_cv2_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for k in ("synthetic_480x640", "picture_684x1054", "frame_224x224"):
        got = O.resize_bilinear_u8(g[k + "_src"], 224, 224)
        assert np.array_equal(got, g[k + "_resized"]), k


def test_resize_edge_cases():
    img = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    assert np.array_equal(O.resize_bilinear_u8(img, 5, 7), img)          # identity
    one = np.full((1, 1, 3), 200, np.uint8)
    assert np.array_equal(O.resize_bilinear_u8(one, 4, 4), np.full((4, 4, 3), 200, np.uint8))  # 1x1 source
    assert O.resize_bilinear_u8(img[:, :, 0], 3, 3).shape == (3, 3)      # single channel


def test_postprocess_matches_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for tag in ("logits", "probs"):
        m = O.postprocess_oracle([g[f"post_{tag}_in"]], (240, 320), 0.5)
        assert np.array_equal(m, g[f"post_{tag}_mask"]), tag
    # strict '>' (src/unet.py:67): a probability exactly at the threshold is background
    p = np.full((1, 1, 4, 4), 0.5, np.float32)
    assert O.postprocess_oracle([p], (4, 4), 0.5).max() == 0
    # 3-D output form (src/unet.py:55-56) and int8 input (src/unet.py:59-60)
    assert O.postprocess_oracle([np.ones((1, 4, 4), np.float32)], (4, 4), 0.5).min() == 255
    assert O.postprocess_oracle([np.full((1, 1, 4, 4), 5, np.int8)], (4, 4), 0.5).min() == 255


def test_bf16_emulation_is_close_to_fp32():
    m = _model([64, 128])
    x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y = m(x)
    ye = O.forward_bf16_emulated(m, x)
    assert (y - ye).abs().max().item() < 0.05 * max(1.0, y.abs().max().item())


def _golden_mod(golden_dir):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ipm_front_end_matches_cv2_goldens(golden_dir):
    """src/unet_ros_node.py:297-313: warpPerspective -> BGR2RGB -> resize, against outputs of cv2 itself (ipm.npz)."""
    g = np.load(os.path.join(golden_dir, "ipm.npz"))
    assert np.array_equal(O.invert3x3(g["M"]), g["Minv"])
    syn = _golden_mod(golden_dir).ipm_synthetic_frame(77)
    warped = O.warp_perspective_u8(syn, g["Minv"], 1055, 685)
    assert np.array_equal(warped[300:364, 400:528], g["syn_warp_crop"])
    assert int(warped.astype(np.int64).sum()) == int(g["syn_warp_sum"])
    got, shape = O.ipm_preprocess_oracle(syn, g["Minv"])
    assert shape == (685, 1055) and np.array_equal(got[0], g["syn_resized"])
    got, _ = O.ipm_preprocess_oracle(g["real_src_half"], g["Minv"])
    assert np.array_equal(got[0], g["real_half_resized"])


def test_resize_up_and_special_cases_match_cv2_goldens(golden_dir):
    """src/unet.py:70 mask up-resize and the shapes cv::resize special-cases (border rows, 2x2 decimation, copy)."""
    g = np.load(os.path.join(golden_dir, "ipm.npz"))
    assert np.array_equal(O.resize_bilinear_u8(g["mask224"], 685, 1055), g["mask_up"])
    for tag in ("up2x", "area2x", "same", "ragged"):
        dst = g[f"gray_{tag}_dst"]
        assert np.array_equal(O.resize_bilinear_u8(g[f"gray_{tag}_src"], dst.shape[0], dst.shape[1]), dst), tag
