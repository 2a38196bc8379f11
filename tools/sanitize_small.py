"""Small end-to-end run for compute-sanitizer: one eval forward, one fused train step, IPM front end (tiny shapes)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402

torch.manual_seed(0)
net = U.UNet(3, 1, [64, 128]).cuda().eval()
frames = torch.randint(0, 256, (3, 40, 56, 3), dtype=torch.uint8).cuda()
_, _, m = net.predict_mask(frames, size=(32, 48), want=("mask",))
net.train()
step = U.FusedTrainStep(net, cuda_graph=False)
x = torch.randn(4, 3, 32, 32).cuda()
y = (torch.rand(4, 1, 32, 32) < 0.2).float().cuda()
print(step.step(x, y).tolist())
M = np.array([[1.1, 0.2, -3.0], [0.05, 0.9, 2.0], [0.0, -0.004, 1.0]])
x4 = U.ops.preprocess_warp_u8(frames, M, (70, 50), (32, 48))
up = U.ops.resize_gray_u8(m, (50, 70))
torch.cuda.synchronize()
print("ok", m.shape, x4.shape, up.shape)
