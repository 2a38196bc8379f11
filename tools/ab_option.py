"""Same-box A/B of a library option (unet_b200_set_option): inference frames/s and training ms/step with the option at 0 and 1.
   python tools/ab_option.py pdl            # values 0 1 0 1
   python tools/ab_option.py <name> A B      # explicit pair of values"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402
from unet_lane_detection_b200._lib import check, lib  # noqa: E402

name = sys.argv[1].encode()
vals = [int(v) for v in sys.argv[2:4]] if len(sys.argv) >= 4 else [0, 1]
for val in vals * 2:
    check(lib.unet_b200_set_option(name, val))
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128, 256, 512]).cuda().eval()
    fr = torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, device="cuda")
    for _ in range(5):
        net.predict_mask(fr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        net.predict_mask(fr)
    e1.record()
    torch.cuda.synchronize()
    fps = 256 * 40 / (e0.elapsed_time(e1) / 1e3)
    del fr
    net._engines.clear()
    torch.cuda.empty_cache()
    net.train()
    step = U.FusedTrainStep(net)
    x = torch.randn(64, 3, 224, 224, device="cuda")
    y = (torch.rand(64, 1, 224, 224, device="cuda") < 0.085).float()
    for _ in range(5):
        step.step(x, y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        step.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    print(f"{sys.argv[1]}={val}: inference {fps:.0f} frames/s, training {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)
    del step, net, x, y
    torch.cuda.empty_cache()
