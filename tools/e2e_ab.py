"""Same-box A/B of a library option on the HOST-buffer entry (UNet.infer_host -> unet_b200_infer_u8_host_stream):
frames/s from pinned host frames of a given size to host masks.
   python tools/e2e_ab.py host_hybrid 480 640          # values 0 1 0 1, batch 256
   python tools/e2e_ab.py pre_bulk 480 640 0 1 256"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402
from unet_lane_detection_b200._lib import check, lib  # noqa: E402

name = sys.argv[1]
hs, ws = int(sys.argv[2]), int(sys.argv[3])
vals = [int(sys.argv[4]), int(sys.argv[5])] if len(sys.argv) >= 6 else [0, 1]
batch = int(sys.argv[6]) if len(sys.argv) >= 7 else 256
frames = torch.randint(0, 256, (batch, hs, ws, 3), dtype=torch.uint8).pin_memory()
mask = torch.empty(batch, 224, 224, dtype=torch.uint8).pin_memory()
for val in vals * 2:
    check(lib.unet_b200_set_option(name.encode(), val))
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128, 256, 512]).cuda().eval()
    net.b200_chunk = 256
    for _ in range(4):
        net.infer_host(frames, mask_out=mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        net.infer_host(frames, mask_out=mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"{name}={val}: {hs}x{ws} host frames, batch {batch}: {ms:.3f} ms/call, {batch / ms * 1e3:.0f} frames/s", flush=True)
    del net
    torch.cuda.empty_cache()
