"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

What it pins:
  * unet_manifest.json   - state_dict keys/shapes/dtypes + parameter counts of the reference's
                           `class UNet` listing (README.md:1418-1481) executed verbatim.
  * unet_small.npz       - logits of that class on seeded inputs for two small configurations,
                           default init + randomised BN (oracle.randomize_bn_), eval mode.
  * loss_small.npz       - BCEDiceLoss (README.md:1855-1893, executed verbatim) values + input grads.
  * preprocess.npz       - cv2.resize outputs on the reference's sample images / a seeded frame,
                           and postprocess masks produced by the reference's own function body
                           (src/unet.py:44-72) for seeded logits.
Weights are NOT stored (31 M parameters); they are regenerated from the same torch seed and the
manifest carries a checksum to detect a drifting initialiser.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import unet_oracle as O  # noqa: E402


def load_listing(first: int, last: int) -> str:
    with open(os.path.join(REF, "README.md"), encoding="utf-8") as f:
        lines = f.readlines()
    return "".join(lines[first - 1:last])


def reference_unet_class():
    ns = {}
    exec(load_listing(1418, 1481), ns)  # import torch ... class UNet (up to `return self.output(x)`)
    return ns["UNet"]


def reference_loss_class():
    ns = {"torch": torch, "nn": torch.nn}
    exec(load_listing(1855, 1893), ns)
    return ns["BCEDiceLoss"]


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


def main():
    import cv2

    RefUNet = reference_unet_class()
    manifest = {}
    for name, feats in (("default", [64, 128, 256, 512]), ("deployed", [32, 64, 128])):
        torch.manual_seed(0)
        m = RefUNet(3, 1, features=feats)
        sd = m.state_dict()
        manifest[name] = {
            "features": feats,
            "n_params": sum(p.numel() for p in m.parameters()),
            "n_entries": len(sd),
            "abs_checksum_seed0": checksum(sd),
            "entries": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
        }
    assert manifest["default"]["n_params"] == 31037633  # README.md:2288
    with open(os.path.join(HERE, "unet_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)

    # ---- logits goldens (small spatial sizes so the CPU suite stays fast)
    out = {}
    for tag, feats, hw, batch in (("f64x2", [64, 128], (32, 48), 2), ("default", [64, 128, 256, 512], (32, 32), 1)):
        torch.manual_seed(0)
        ref = RefUNet(3, 1, features=feats).eval()
        O.randomize_bn_(ref, seed=1)
        O.scale_head_(ref, 40.0)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(batch, 3, *hw, generator=g)
        with torch.no_grad():
            y = ref(x)
        # the restatement must agree with the listing bit-for-bit on the same weights
        mine = O.UNetOracle(3, 1, feats).eval()
        mine.load_state_dict(ref.state_dict())
        with torch.no_grad():
            y2 = mine(x)
        assert torch.equal(y, y2), "oracle restatement differs from the reference listing"
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_logits"] = y.numpy()
        out[f"{tag}_checksum"] = np.float64(checksum(ref.state_dict()))
    np.savez_compressed(os.path.join(HERE, "unet_small.npz"), **out)

    # ---- loss goldens
    RefLoss = reference_loss_class()
    g = torch.Generator().manual_seed(7)
    logits = (torch.randn(2, 1, 32, 32, generator=g) * 2).requires_grad_(True)
    target = (torch.rand(2, 1, 32, 32, generator=g) < 0.085).float()
    crit = RefLoss(0.5, 0.5, pos_weight=torch.tensor([3.0]))
    total, bce, dice = crit(logits, target)
    total.backward()
    mine = O.BCEDiceLossOracle(0.5, 0.5, pos_weight=torch.tensor([3.0]))
    t2, b2, d2 = mine(logits.detach(), target)
    assert torch.allclose(total, t2) and torch.allclose(bce, b2) and torch.allclose(dice, d2)
    np.savez_compressed(os.path.join(HERE, "loss_small.npz"), logits=logits.detach().numpy(), target=target.numpy(),
                        total=total.item(), bce=bce.item(), dice=dice.item(), grad=logits.grad.numpy())

    # ---- preprocess / postprocess goldens
    pre = {}
    rng = np.random.default_rng(1234)
    frames = {
        "synthetic_480x640": rng.integers(0, 256, (480, 640, 3), dtype=np.uint8),
        "picture_684x1054": cv2.imread(os.path.join(REF, "picture.jpg")),
        "frame_224x224": cv2.imread(os.path.join(REF, "test_images", "frame_001410.jpg")),
    }
    for k, img in frames.items():
        r = cv2.resize(img, (224, 224))  # src/unet.py:33
        assert np.array_equal(r, O.resize_bilinear_u8(img, 224, 224)), k
        if k.startswith("synthetic"):
            pre[k + "_src"] = img
            pre[k + "_resized"] = r
        else:
            # keep the fixture small; never below the model size (cv2's up-scale path is not modelled)
            small = cv2.resize(img, (img.shape[1] // 2, img.shape[0] // 2)) if min(img.shape[:2]) >= 448 else img
            pre[k + "_src"] = small
            pre[k + "_resized"] = cv2.resize(small, (224, 224))
            assert np.array_equal(pre[k + "_resized"], O.resize_bilinear_u8(small, 224, 224))
    # the reference's postprocess body, executed verbatim (needs only np/cv2)
    with open(os.path.join(REF, "src", "unet.py"), encoding="utf-8") as f:
        src_lines = f.readlines()
    body = "".join(src_lines[43:72])  # def postprocess_output ... return binary_mask
    ns = {"np": np, "cv2": cv2}
    exec("class _R:\n" + body, ns)
    ref_post = ns["_R"]().postprocess_output
    for tag, arr in (("logits", (rng.standard_normal((1, 1, 224, 224)) * 3).astype(np.float32)),
                     ("probs", rng.random((1, 1, 224, 224)).astype(np.float32))):
        m = ref_post([arr], (240, 320), 0.5)
        assert np.array_equal(m, O.postprocess_oracle([arr], (240, 320), 0.5)), tag
        pre[f"post_{tag}_in"] = arr
        pre[f"post_{tag}_mask"] = m
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **pre)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
