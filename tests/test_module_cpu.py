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
    m.train()                                                  # the training path has no CPU fallback either
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 32, 32))
    with pytest.raises(ValueError):
        U.ops.conv3x3(torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16), torch.zeros(64, 9, 64), torch.zeros(64))


def test_flat_parameters_keep_identity_and_state_dict():
    """training.flatten_parameters_: every Parameter becomes a view of one flat fp32 buffer (parameters() order) without
    changing identity, values, names or the state_dict - the optimizer created before flattening stays valid."""
    from unet_lane_detection_b200.training import flatten_parameters_
    torch.manual_seed(0)
    m = U.UNet(3, 1, [64, 128])
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    ids = [id(p) for p in m.parameters()]
    flat = flatten_parameters_(m)
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    assert [id(p) for p in m.parameters()] == ids
    off = 0
    for p in m.parameters():
        assert p.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert flatten_parameters_(m) is flat                      # idempotent
    flat.add_(1.0)                                             # a kernel writing the flat buffer updates every parameter
    assert torch.equal(m.output.bias.detach(), before["output.bias"] + 1.0)
    assert {id(p) for g in opt.param_groups for p in g["params"]} == set(ids)


def test_trainer_layout_matches_parameters():
    """The C library's flat parameter layout (unet_b200_trainer_tensor_offset) is model.parameters() order."""
    import ctypes as C
    from unet_lane_detection_b200._lib import check, lib
    for feats in ([64, 128], [64, 128, 256, 512]):
        m = U.UNet(3, 1, feats)
        h = C.c_void_p()
        check(lib.unet_b200_trainer_create(C.byref(h), 8, 64, 64, 3, 1, (C.c_int * len(feats))(*feats), len(feats)))
        nt = lib.unet_b200_trainer_num_tensors(h)
        offs = [lib.unet_b200_trainer_tensor_offset(h, i) for i in range(nt + 1)]
        sizes = [p.numel() for p in m.parameters()]
        assert nt == len(sizes) and [b - a for a, b in zip(offs, offs[1:])] == sizes
        assert lib.unet_b200_trainer_num_params(h) == sum(sizes)
        assert lib.unet_b200_trainer_workspace_bytes(h) > 0
        lib.unet_b200_trainer_destroy(h)
    h = C.c_void_p()
    assert lib.unet_b200_trainer_create(C.byref(h), 8, 64, 64, 3, 1, (C.c_int * 2)(64, 192), 2) != 0   # not a power of two
    assert b"features" in lib.unet_b200_last_error()


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


def test_deployed_topology_key_remap_round_trip():
    """checkpoints of the deployed [32,64,128] topology (block names inc/down*/up*/conv*/outc as found
    inside model/lane_unet*.rknn) are remapped positionally onto the reference listing's keys."""
    from unet_lane_detection_b200.executor import _features_from_state_dict, remap_milesial_state_dict
    torch.manual_seed(0)
    ref = O.UNetOracle(3, 1, [32, 64, 128])
    assert sum(p.numel() for p in ref.parameters()) == 1927009
    sd = ref.state_dict()
    names = {"encoder_blocks.0": "inc", "encoder_blocks.1": "down1", "encoder_blocks.2": "down2", "bottleneck": "down3",
             "decoder_blocks.0": "up1", "decoder_blocks.1": "conv1", "decoder_blocks.2": "up2", "decoder_blocks.3": "conv2",
             "decoder_blocks.4": "up3", "decoder_blocks.5": "conv3", "output": "outc"}
    inner = {"0": "double_conv.0", "1": "double_conv.1", "3": "double_conv.3", "4": "double_conv.4"}
    renamed = {}
    for k, v in sd.items():
        pre = next(p for p in sorted(names, key=len, reverse=True) if k.startswith(p + "."))
        rest = k[len(pre) + 1:].split(".")
        if rest[0] in inner and len(rest) > 1:
            rest[0] = inner[rest[0]]
        renamed[names[pre] + "." + ".".join(rest)] = v
    assert not any(k.startswith("encoder_blocks") for k in renamed)
    back = remap_milesial_state_dict(renamed)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    assert _features_from_state_dict(back) == ([32, 64, 128], 3, 1)
    m = U.UNet(3, 1, [32, 64, 128])
    m.load_state_dict(back)
    assert remap_milesial_state_dict(sd) is sd
    with pytest.raises(ValueError):
        remap_milesial_state_dict({"foo.weight": torch.zeros(1)})


def test_weights_key_and_copies_of_the_module():
    """The per-call staleness key of a live module (identity + in-place version of all 118 tensors, read through cached
    (dict, name) slots) notices in-place updates, load_state_dict and a replaced Parameter object; copy.deepcopy / torch.save
    of the module leave the plans and other runtime state behind (a copy builds its own)."""
    import copy
    import io
    net = U.UNet(3, 1, [64, 128]).eval()
    k0 = net._weights_key()
    assert len(k0) == 1 + len(list(net.parameters())) + len(list(net.buffers())) and net._weights_key() == k0
    with torch.no_grad():
        net.output.bias.add_(1.0)
    k1 = net._weights_key()
    assert k1 != k0
    net.load_state_dict(copy.deepcopy(net.state_dict()))
    k2 = net._weights_key()
    assert k2 != k1
    net.output.weight = torch.nn.Parameter(torch.zeros_like(net.output.weight))
    assert net._weights_key() != k2
    net._engines["stand-in"] = object()
    twin = copy.deepcopy(net)
    assert len(twin._engines) == 0 and twin._last_engine is None and "_b200_key_slots" not in twin.__dict__
    assert twin._weights_key() != net._weights_key()          # its own tensors
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), twin.state_dict().values()))
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert len(back._engines) == 0 and back.features == net.features and back.b200_chunk == net.b200_chunk
