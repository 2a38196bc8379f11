"""CPU: the C-ABI library loads, exports every symbol include/unet_b200.h declares, and rejects bad
arguments without needing a GPU."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "unet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(unet_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    from unet_lane_detection_b200 import _lib
    declared = _declared()
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\b(unet_b200_\w+)\b", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/unet_b200.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == declared, "ctypes binding table and header disagree"


def test_header_compiles_as_c():
    src = '#include "unet_b200.h"\nint main(void){return UB_OK;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_argument_validation_without_gpu():
    from unet_lane_detection_b200._lib import lib
    h = C.c_void_p()
    feats = (C.c_int * 4)(64, 128, 256, 512)
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 1, feats, 4) == 0
    assert lib.unet_b200_plan_num_convs(h) == 18
    assert lib.unet_b200_plan_num_layers(h) == 23       # 18 conv3x3 (stem included) + 4 convT + head
    assert lib.unet_b200_plan_workspace_bytes(h) > 0 and lib.unet_b200_plan_weight_bytes(h) > 62_000_000
    # forward before bind -> state error, never a crash
    assert lib.unet_b200_forward(h, C.c_void_p(8), 1, None, None, None, 0.5, None) == -4
    lib.unet_b200_plan_destroy(h)
    odd = (C.c_int * 2)(64, 100)                      # any positive widths (stored zero-extended to multiples of 64)
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 1, odd, 2) == 0
    lib.unet_b200_plan_destroy(h)
    bad = (C.c_int * 2)(64, 0)
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 1, bad, 2) == -1
    assert b"features[1]" in lib.unet_b200_last_error()
    wide = (C.c_int * 2)(320, 640)                    # the first block's width is bounded by the stem kernel
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 1, wide, 2) == -1
    assert lib.unet_b200_plan_create(C.byref(h), 8, 100, 224, 3, 1, feats, 4) == -1   # H not divisible by 16
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 5, 1, feats, 4) == -1   # in_channels > 4
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 0, feats, 4) == -1   # out_channels < 1
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 3, feats, 4) == 0    # any out_channels (README.md:1447)
    lib.unet_b200_plan_destroy(h)
    assert lib.unet_b200_conv3x3(None, 64, None, 0, None, None, 1, 8, 8, 64, 1, None, None, None) == -1


def test_layer_table_matches_survey_appendix_b():
    """The plan's layer table must reproduce the per-layer GEMM shapes of SURVEY.md Appendix B (73.756 GFLOP/frame)."""
    from unet_lane_detection_b200._lib import lib
    h = C.c_void_p()
    feats = (C.c_int * 4)(64, 128, 256, 512)
    assert lib.unet_b200_plan_create(C.byref(h), 1, 224, 224, 3, 1, feats, 4) == 0
    flops = 0.0
    for i in range(lib.unet_b200_plan_num_layers(h)):
        info = (C.c_int * 8)()
        assert lib.unet_b200_plan_layer_info(h, i, info) == 0
        kind, H, W, cin, cout, taps, _bn, _pool = list(info)
        cin = 3 if kind == 0 else cin
        flops += 2.0 * H * W * cout * (4 if kind == 2 else 1) * taps * cin
    lib.unet_b200_plan_destroy(h)
    assert abs(flops / 1e9 - 73.756) < 0.01


def test_workspace_is_shared_by_liveness():
    """layer outputs share the workspace by liveness. Default network: 19.3 MB per frame of capacity (the
    live set at dec3.conv0: skip + up-sampled tensor + output) instead of 64.1 MB; <= 5.5 GB at batch 256."""
    from unet_lane_detection_b200._lib import lib
    h = C.c_void_p()
    feats = (C.c_int * 4)(64, 128, 256, 512)
    assert lib.unet_b200_plan_create(C.byref(h), 256, 224, 224, 3, 1, feats, 4) == 0
    ws, un = lib.unet_b200_plan_workspace_bytes(h), lib.unet_b200_plan_workspace_unshared_bytes(h)
    lib.unet_b200_plan_destroy(h)
    assert ws <= 5.5e9, ws
    assert ws / 256 <= 20.0e6 and un / 256 >= 57e6, (ws / 256, un / 256)   # (unshared: 64.1 MB minus the never-written last output)
    # the unfused head also writes the last activation; it fits the space the skip / up-sampled tensors just left
    assert lib.unet_b200_set_option(b"fuse_head", 0) == 0
    try:
        assert lib.unet_b200_plan_create(C.byref(h), 256, 224, 224, 3, 1, feats, 4) == 0
        ws2 = lib.unet_b200_plan_workspace_bytes(h)
        lib.unet_b200_plan_destroy(h)
    finally:
        lib.unet_b200_set_option(b"fuse_head", 1)
    assert ws <= ws2 <= 26.5e6 * 256


def test_backward_stages_cover_the_flat_gradient_once():
    """Stage ranges of the staged backward (trainer_stage_range) tile [0, n_params) exactly; the bucket plan built from them
    keeps that property and ends every bucket on a 4-element boundary (16-byte accesses of the exchange kernels)."""
    from unet_lane_detection_b200._lib import lib
    from unet_lane_detection_b200.training import plan_buckets, shard_of
    for feats_l, hw in (([64, 128, 256, 512], 224), ([64, 128], 32), ([128, 256, 512], 64)):
        h = C.c_void_p()
        feats = (C.c_int * len(feats_l))(*feats_l)
        assert lib.unet_b200_trainer_create(C.byref(h), 64, hw, hw, 3, 1, feats, len(feats_l)) == 0
        n = lib.unet_b200_trainer_num_params(h)
        S = lib.unet_b200_trainer_num_stages(h)
        assert S == 2 * len(feats_l) + 2
        rs = []
        for s in range(S):
            lo, hi = C.c_longlong(), C.c_longlong()
            assert lib.unet_b200_trainer_stage_range(h, s, C.byref(lo), C.byref(hi)) == 0
            rs.append((lo.value, hi.value))
        lib.unet_b200_trainer_destroy(h)
        cover = sorted(rs)
        assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        for mb in (1, 20000, 2 << 20, 1 << 40):
            bk = plan_buckets(rs, mb, min(1 << 16, mb))
            iv = sorted((lo, hi) for _, lo, hi in bk)
            assert iv[0][0] == 0 and iv[-1][1] == n and all(a[1] == b[0] for a, b in zip(iv, iv[1:])), (feats_l, mb, bk)
            assert [s for s, _, _ in bk] == sorted(s for s, _, _ in bk)
            for stage, lo, hi in bk:       # a bucket is sent only after every stage that contributes to it
                assert all(s <= stage for s, (a, b) in enumerate(rs) if a < hi and b > lo)
                assert lo % 4 == 0
                for world in (2, 8):
                    parts = [shard_of(lo, hi, r, world) for r in range(world)]
                    assert parts[0][0] == lo and parts[-1][1] == hi and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
                    assert all(a % 4 == 0 for a, b in parts if b > a)      # (empty parts are never launched)
    # default network: five buckets, the big decoder / bottleneck ones leave mid-backward
    assert len(plan_buckets(rs_default())) == 5


def rs_default():
    from unet_lane_detection_b200._lib import lib
    h = C.c_void_p()
    feats = (C.c_int * 4)(64, 128, 256, 512)
    assert lib.unet_b200_trainer_create(C.byref(h), 64, 224, 224, 3, 1, feats, 4) == 0
    rs = []
    for s in range(lib.unet_b200_trainer_num_stages(h)):
        lo, hi = C.c_longlong(), C.c_longlong()
        lib.unet_b200_trainer_stage_range(h, s, C.byref(lo), C.byref(hi))
        rs.append((lo.value, hi.value))
    lib.unet_b200_trainer_destroy(h)
    return rs


def test_host_pipeline_schedule_without_gpu():
    """Pass / piece schedule of unet_b200_infer_u8_host_stream (pure host logic, include/unet_b200.h): a pass of 256 frames is
    cut into the geometric pieces 16 + 32 + 64 + 144; the preprocess, the stem, enc0.conv1 and the fused-head conv run once
    per piece, every other layer once per pass. Camera-size sources (copy-bound) get a short first pass (64 = 16 + 48) and
    then 192 = 16 + 32 + 64 + 80; 600 frames are 256 + 256 + 88 (16 + 32 + 40)."""
    from unet_lane_detection_b200._lib import check, lib
    h = C.c_void_p()
    feats = (C.c_int * 4)(64, 128, 256, 512)
    assert lib.unet_b200_plan_create(C.byref(h), 256, 224, 224, 3, 1, feats, 4) == 0
    per_pass = lib.unet_b200_forward_launches(h) + 1          # + preprocess
    assert per_pass == 23 and lib.unet_b200_plan_host_pieces(h) == 4
    assert lib.unet_b200_infer_stream_launches(h, 256, 224, 224) == per_pass + 4 * 3
    assert lib.unet_b200_infer_stream_launches(h, 256, 480, 640) == (per_pass + 4 * 1) + (per_pass + 4 * 3)
    assert lib.unet_b200_infer_stream_launches(h, 600, 224, 224) == 2 * (per_pass + 4 * 3) + (per_pass + 4 * 2)
    assert lib.unet_b200_infer_stream_launches(h, 1, 224, 224) == per_pass
    lib.unet_b200_plan_destroy(h)
    try:    # equal pieces (host_geometric = 0): eight pieces of 32; plans copy the defaults when they are created
        check(lib.unet_b200_set_option(b"host_geometric", 0))
        assert lib.unet_b200_plan_create(C.byref(h), 256, 224, 224, 3, 1, feats, 4) == 0
        assert lib.unet_b200_plan_host_pieces(h) == 8
        assert lib.unet_b200_infer_stream_launches(h, 256, 224, 224) == per_pass + 4 * 7
        lib.unet_b200_plan_destroy(h)
        check(lib.unet_b200_set_option(b"host_pieces", 0))
        assert lib.unet_b200_plan_create(C.byref(h), 256, 224, 224, 3, 1, feats, 4) == 0
        assert lib.unet_b200_plan_host_pieces(h) == 0
        assert lib.unet_b200_infer_stream_launches(h, 256, 224, 224) == 2 * per_pass     # pass-granular: 64 + 192
        lib.unet_b200_plan_destroy(h)
    finally:
        check(lib.unet_b200_set_option(b"host_geometric", 1))
        check(lib.unet_b200_set_option(b"host_pieces", 8))
