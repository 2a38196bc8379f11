"""Same-box A/B of a plan-level library option on small and mid-size plans: device-resident ms per UNet.predict_mask call.
    python tools/ab_plan_option.py small_n 1 2      (1 = wave model, the default; 2 = narrower blocks only while a layer cannot
                                                     give every SM a tile)
    python tools/ab_plan_option.py pdl 0 1
profiles/r2_ab_small_n.log was taken when small_n's two values had the opposite meaning (there 2 = wave model; its first six
lines with a 128-wide tile cost of 0.5, the rest with the calibrated 0.56)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402
from unet_lane_detection_b200._lib import check, lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "small_n"
v0, v1 = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 2)
cases = [(1, 224, 224), (2, 224, 224), (4, 224, 224), (8, 224, 224), (16, 224, 224), (32, 224, 224), (64, 224, 224), (1, 480, 640), (4, 480, 640), (16, 480, 640)]
for B, H, W in cases:
    res = []
    for val in (v0, v1, v0, v1):
        check(lib.unet_b200_set_option(name.encode(), val))
        torch.manual_seed(0)
        net = U.UNet(3, 1, [64, 128, 256, 512]).cuda().eval()
        net.b200_chunk = B
        fr = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
        for _ in range(5):
            net.predict_mask(fr, size=(H, W))
        torch.cuda.synchronize()
        n = max(10, min(200, 4096 // B))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            net.predict_mask(fr, size=(H, W))
        e1.record()
        torch.cuda.synchronize()
        res.append(f"{name}={val}: {e0.elapsed_time(e1) / n:.3f} ms")
        del net, fr
        torch.cuda.empty_cache()
    print(f"batch {B} @ {H}x{W}: " + ", ".join(res), flush=True)
