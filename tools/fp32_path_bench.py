"""Throughput of the fp32-class eval forward (b200_precision = 'fp32') next to the bf16 path, same weights and input.
Usage (on a B200): python tools/fp32_path_bench.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
net = U.UNet(3, 1, [64, 128, 256, 512]).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
out = {}
for prec in ("bf16", "fp32"):
    net.b200_precision = prec
    with torch.no_grad():
        for _ in range(3):
            y = net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = net(x)
        e1.record()
        torch.cuda.synchronize()
    out[prec] = (B * 10 / (e0.elapsed_time(e1) / 1e3), y.float().cpu())
print(f"batch {B}: bf16 {out['bf16'][0]:.0f} frames/s, fp32-class {out['fp32'][0]:.0f} frames/s "
      f"({out['fp32'][0] * 3 * 73.756e9 / 1e12:.0f} TFLOP/s of bf16 tensor work), "
      f"max|bf16 - fp32| = {(out['bf16'][1] - out['fp32'][1]).abs().max().item():.3e}")
