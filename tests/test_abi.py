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
    bad = (C.c_int * 2)(64, 100)
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 1, bad, 2) == -1
    assert b"multiple of 32" in lib.unet_b200_last_error()
    assert lib.unet_b200_plan_create(C.byref(h), 8, 100, 224, 3, 1, feats, 4) == -1   # H not divisible by 16
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 5, 1, feats, 4) == -1   # in_channels > 4
    assert lib.unet_b200_plan_create(C.byref(h), 8, 224, 224, 3, 2, feats, 4) == -1   # out_channels != 1
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
