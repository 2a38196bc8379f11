"""unet-lane-detection_b200: B200 (sm_100a) U-Net hot path behind the reference's interfaces.

    from unet_lane_detection_b200 import UNet, B200_model_container

`UNet` mirrors the reference's nn.Module (README.md:1421-1481); `B200_model_container` mirrors the
executor plugin API (src/py_utils/rknn_executor.py). Both drive libunet_b200.so (hand-written CUDA,
C ABI in include/unet_b200.h) and raise if it is missing - there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library has not been built)
from .ops import (  # noqa: F401
    conv3x3, convT2x2, head, maxpool2x2, nchw_to_nhwc4, pack_conv3x3, pack_convT2x2, pack_stem, pack_stem_tc, preprocess_u8,
    preprocess_warp_u8, resize_gray_u8, stem_conv, stem_conv_tc, warp_perspective_u8,
)
from .unet import UNet  # noqa: F401
from .executor import B200_model_container, B200LaneInference, B200LanePipeline  # noqa: F401
from .training import FusedTrainStep, bce_dice_loss  # noqa: F401
from .loop import cosine_warm_restarts_lr, fit, train_one_epoch, validate, validation_metrics  # noqa: F401
from .torchscript import LaneGraph, export_torchscript  # noqa: F401  (importing it registers the unet_b200::infer operator)

__all__ = ["UNet", "B200_model_container", "B200LaneInference", "B200LanePipeline", "FusedTrainStep", "bce_dice_loss", "fit", "train_one_epoch",
           "validate", "validation_metrics", "cosine_warm_restarts_lr", "export_torchscript", "LaneGraph"]
