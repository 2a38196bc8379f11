"""Preprocess kernels alone: GB/s of the same-size copy path and of the cv2-exact tile kernel on camera frames.
    python tools/pre_bench.py [frames]      (also the target of the ncu capture in profiles/r2_ncu_full_preprocess.txt)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {}


def rows_touched(hs, h):
    """Distinct source rows cv2's bilinear taps read for h output rows (down-scaling by more than 2 skips rows)."""
    import numpy as np
    f = ((np.arange(h) + 0.5) * (hs / h) - 0.5).astype(np.float32)
    s0 = np.floor(f).astype(np.int64)
    return len(set(np.clip(s0, 0, hs - 1)) | set(np.clip(s0 + 1, 0, hs - 1)))



for tag, hs, ws in (("same_224x224", 224, 224), ("camera_480x640", 480, 640), ("bev_685x1055", 685, 1055), ("hd_960x1280", 960, 1280)):
    f = torch.randint(0, 256, (n, hs, ws, 3), dtype=torch.uint8, device="cuda")
    y = torch.empty(n, 224, 224, 4, dtype=torch.bfloat16, device="cuda")
    from unet_lane_detection_b200._lib import check, f3, lib
    from unet_lane_detection_b200.ops import MEAN_255, STD_255
    st = torch.cuda.current_stream().cuda_stream

    def run():
        check(lib.unet_b200_preprocess_u8(f.data_ptr(), n, hs, ws, ws * 3, hs * ws * 3, 224, 224, 1, f3(MEAN_255), f3(STD_255),
                                          y.data_ptr(), None, st))

    run()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    # bytes the kernel has to move: the source ROWS its taps touch (whole rows: every 32-byte sector of a touched row holds a
    # tap for horizontal scales < 10) + the NHWC4 bf16 output
    rd, wr = n * rows_touched(hs, 224) * ws * 3, n * 224 * 224 * 8
    out[tag] = {"frames": n, "us": best * 1e3, "read_mb": rd / 1e6, "frame_mb": n * hs * ws * 3 / 1e6, "write_mb": wr / 1e6,
                "gbs": (rd + wr) / (best / 1e3) / 1e9}
    del f, y
# IPM front end (warp_preprocess_u8_kernel): camera frames -> bird's-eye 1055x685 (never stored) -> 224x224 network input
import numpy as np  # noqa: E402
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ipm.npz"))
f = torch.randint(0, 256, (n, 480, 640, 3), dtype=torch.uint8, device="cuda")
U.ops.preprocess_warp_u8(f, g["M"], (1055, 685), (224, 224))
torch.cuda.synchronize()
best = 1e30
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        U.ops.preprocess_warp_u8(f, g["M"], (1055, 685), (224, 224))
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
out["ipm_480x640_to_1055x685_to_224"] = {"frames": n, "us": best * 1e3, "frame_mb": n * 480 * 640 * 3 / 1e6, "write_mb": n * 224 * 224 * 8 / 1e6,
                                         "gbs_whole_frames": (n * 480 * 640 * 3 + n * 224 * 224 * 8) / (best / 1e3) / 1e9}
print(json.dumps(out))
