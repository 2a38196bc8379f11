"""CPU: the drop-in nn.Module / executor boundary (no kernels run here)."""
import json
import os

import numpy as np
import pytest
import torch

import unet_lane_detection_b200 as U
from oracle import unet_oracle as O


def test_state_dict_layout_is_the_reference_layout(golden_dir):
    man = json.load(open(os.path.join(golden_dir, "unet_manifest.json")))
    for name in ("default", "deployed"):
        m = U.UNet(3, 1, man[name]["features"])
        assert [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()] == man[name]["entries"]


def test_constructor_signature_and_defaults():
    import inspect
    sig = inspect.signature(U.UNet.__init__)
    assert list(sig.parameters)[1:] == ["in_channels", "out_channels", "features"]
    assert sig.parameters["in_channels"].default == 3 and sig.parameters["out_channels"].default == 1
    assert sig.parameters["features"].default == [64, 128, 256, 512]


def test_checkpoint_interchange_with_reference_forms(tmp_path):
    ref = O.UNetOracle(3, 1, [64, 128])
    m = U.UNet(3, 1, [64, 128])
    m.load_state_dict(ref.state_dict())                       # bare state_dict (README.md:2231)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), ref.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    ref.load_state_dict(m.state_dict())                       # and back
    path = tmp_path / "best_model.pth"                        # dict form (README.md:2205-2214)
    torch.save({"epoch": 3, "model_state_dict": ref.state_dict(), "best_dice": 0.9}, path)
    from unet_lane_detection_b200.executor import _features_from_state_dict
    feats, cin, cout = _features_from_state_dict(torch.load(path, weights_only=True)["model_state_dict"])
    assert (feats, cin, cout) == ([64, 128], 3, 1)
    assert list(dict(m.named_parameters())) == list(dict(ref.named_parameters()))   # optimizer sees the same params


def test_no_cpu_fallback():
    m = U.UNet(3, 1, [64, 128]).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 32, 32))
    m.train()
    with pytest.raises(RuntimeError, match="training-mode"):
        m(torch.randn(1, 3, 32, 32))
    with pytest.raises(ValueError):
        U.ops.conv3x3(torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16), torch.zeros(64, 9, 64), torch.zeros(64))


def test_executor_rejects_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        U.B200_model_container(U.UNet(3, 1, [64, 128]))


def test_shard_ranges_cover_batch():
    from unet_lane_detection_b200.parallel import shard_range
    for n in (0, 1, 7, 256, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
