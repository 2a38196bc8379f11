"""ctypes binding of libunet_b200.so (include/unet_b200.h). No torch types cross this boundary:
tensors are passed as raw device pointers plus sizes, the stream as a void*.

The product path has no fallback: if the library is missing this module raises at import.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunet_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python unet-lane-detection_b200/build.py` "
        "(nvcc, sm_100a). There is no CPU / PyTorch fallback for the B200 U-Net path."
    )

lib = C.CDLL(LIB_PATH)

vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

_SIGS = {
    "unet_b200_last_error": (C.c_char_p, []),
    "unet_b200_version": (i32, []),
    "unet_b200_device_ok": (i32, []),
    "unet_b200_plan_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, C.POINTER(i32), i32]),
    "unet_b200_plan_create_ex": (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, C.POINTER(i32), i32, i32]),
    "unet_b200_plan_precision": (i32, [vp]),
    "unet_b200_plan_destroy": (None, [vp]),
    "unet_b200_plan_workspace_bytes": (sz, [vp]),
    "unet_b200_plan_workspace_unshared_bytes": (sz, [vp]),
    "unet_b200_plan_weight_bytes": (sz, [vp]),
    "unet_b200_plan_bind": (i32, [vp, vp, vp]),
    "unet_b200_plan_num_convs": (i32, [vp]),
    "unet_b200_plan_set_conv": (i32, [vp, i32, vp, vp, vp, vp, vp, f32, vp]),
    "unet_b200_plan_set_convT": (i32, [vp, i32, vp, vp, vp]),
    "unet_b200_plan_set_head": (i32, [vp, vp, vp, vp]),
    "unet_b200_forward": (i32, [vp, vp, i32, vp, vp, vp, f32, vp]),
    "unet_b200_forward_launches": (i32, [vp]),
    "unet_b200_forward_profile": (i32, [vp, vp, i32, vp, vp, vp, f32, vp, C.POINTER(f32), i32]),
    "unet_b200_plan_num_layers": (i32, [vp]),
    "unet_b200_plan_layer_info": (i32, [vp, i32, C.POINTER(i32)]),
    "unet_b200_set_option": (i32, [C.c_char_p, i32]),
    "unet_b200_nchw_to_nhwc4": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_nchw_to_nhwc4_f32": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_preprocess_u8_f32": (i32, [vp, i32, i32, i32, sz, sz, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), vp, vp, vp]),
    "unet_b200_preprocess_u8": (i32, [vp, i32, i32, i32, sz, sz, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), vp, vp, vp]),
    "unet_b200_preprocess_warp_u8": (i32, [vp, i32, i32, i32, sz, sz, C.POINTER(C.c_double), i32, i32, i32, i32, i32,
                                           C.POINTER(f32), C.POINTER(f32), vp, vp, vp, vp]),
    "unet_b200_resize_gray_u8": (i32, [vp, i32, i32, i32, vp, i32, i32, vp]),
    "unet_b200_infer_staging_bytes": (sz, [vp, i32, i32]),
    "unet_b200_infer_u8_host": (i32, [vp, vp, vp, i32, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), f32, vp, vp, vp, vp]),
    "unet_b200_infer_stream_staging_bytes": (sz, [vp, i32, i32]),
    "unet_b200_plan_host_pieces": (i32, [vp]),
    "unet_b200_infer_stream_launches": (i32, [vp, i32, i32, i32]),
    "unet_b200_infer_u8_host_stream": (i32, [vp, vp, vp, i32, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), f32, vp, vp, vp, vp]),
    "unet_b200_conv3x3": (i32, [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "unet_b200_convT2x2": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_pack_conv3x3": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp]),
    "unet_b200_pack_convT2x2": (i32, [vp, i32, i32, vp, vp]),
    "unet_b200_pack_stem": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp]),
    "unet_b200_stem_conv": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "unet_b200_pack_stem_tc": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp]),
    "unet_b200_stem_conv_tc": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_head": (i32, [vp, vp, f32, sz, i32, vp, vp, vp, f32, vp]),
    "unet_b200_maxpool2x2": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_conv3x3_split": (i32, [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "unet_b200_convT2x2_split": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_pack_conv3x3_split": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, vp, vp]),
    "unet_b200_pack_convT2x2_split": (i32, [vp, i32, i32, vp, vp]),
    "unet_b200_pack_stem_fp32": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp]),
    "unet_b200_stem_conv_split": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "unet_b200_maxpool2x2_split": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_trainer_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, C.POINTER(i32), i32]),
    "unet_b200_trainer_destroy": (None, [vp]),
    "unet_b200_trainer_workspace_bytes": (sz, [vp]),
    "unet_b200_trainer_num_params": (C.c_longlong, [vp]),
    "unet_b200_trainer_num_tensors": (i32, [vp]),
    "unet_b200_trainer_tensor_offset": (C.c_longlong, [vp, i32]),
    "unet_b200_trainer_bind": (i32, [vp, vp]),
    "unet_b200_train_forward": (i32, [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), f32, f32, vp, vp]),
    "unet_b200_train_backward": (i32, [vp, vp, vp, vp, vp]),
    "unet_b200_train_backward_p2p": (i32, [vp, vp, vp, vp, vp, i32, vp]),
    "unet_b200_adamw_step_p2p": (i32, [vp, vp, i32, i32, vp, vp, vp, C.c_longlong, f32, f32, f32, f32, f32, vp, f32, vp]),
    "unet_b200_adamw_step_multimem": (i32, [vp, vp, vp, i32, i32, vp, vp, C.c_longlong, f32, f32, f32, f32, f32, vp, f32, vp]),
    "unet_b200_multimem_reduce": (i32, [vp, C.c_longlong, C.c_longlong, vp, vp]),
    "unet_b200_bce_dice_loss": (i32, [vp, vp, sz, f32, f32, f32, f32, vp, vp, vp, vp]),
    "unet_b200_validation_metrics": (i32, [vp, vp, sz, f32, f32, f32, f32, f32, vp, vp, vp]),
    "unet_b200_adamw_step": (i32, [vp, vp, vp, vp, sz, f32, f32, f32, f32, f32, i32, f32, vp]),
    "unet_b200_adamw_step_dev": (i32, [vp, vp, vp, vp, sz, f32, vp, f32, f32, f32, f32, vp, f32, vp]),
    "unet_b200_trainer_num_stages": (i32, [vp]),
    "unet_b200_trainer_stage_range": (i32, [vp, i32, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "unet_b200_train_backward_stage": (i32, [vp, i32, vp, vp, vp, vp]),
    "unet_b200_trainer_join": (i32, [vp, vp, vp]),
    "unet_b200_adamw_range_p2p": (i32, [vp, vp, i32, i32, vp, C.c_longlong, C.c_longlong, vp, vp, f32, vp, f32, f32, f32, f32, vp, f32, vp]),
    "unet_b200_adamw_range_multimem": (i32, [vp, vp, vp, C.c_longlong, C.c_longlong, vp, vp, f32, vp, f32, f32, f32, f32, vp, f32, vp]),
    "unet_b200_pack_conv3x3_dgrad": (i32, [vp, i32, i32, vp, vp]),
    "unet_b200_pack_convT2x2_dgrad": (i32, [vp, i32, i32, vp, vp]),
    "unet_b200_conv3x3_wgrad": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, vp]),
    "unet_b200_stem_wgrad": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "unet_b200_convT2x2_wgrad": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "unet_b200_convT2x2_dgrad": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    "unet_b200_bn_relu_train_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "unet_b200_bn_relu_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    "unet_b200_maxpool2x2_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTS = tuple(_SIGS.keys())


class UnetB200Error(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise UnetB200Error(f"libunet_b200 error {rc}: {lib.unet_b200_last_error().decode()}")


# UB_OPTIONS="halo2=0,wgrad2=0": kernel-selection switches (include/unet_b200.h unet_b200_set_option) for A/B runs of
# unmodified scripts. Unknown names fail loudly.
for _kv in filter(None, os.environ.get("UB_OPTIONS", "").split(",")):
    _k, _, _v = _kv.partition("=")
    check(lib.unet_b200_set_option(_k.strip().encode(), int(_v)))


def f3(vals):
    return (f32 * 3)(*[float(v) for v in vals])
