"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): one eval forward per plan kind, the preprocess
kernels (copy path, tile kernel in span and sparse mode, fp32 output), the multi-channel head, a fused train step of the
default-style and of the zero-extended ([32,64]) topology through the staged backward, and the IPM front end - tiny shapes.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402
from unet_lane_detection_b200._lib import check, lib  # noqa: E402

torch.manual_seed(0)
net = U.UNet(3, 1, [64, 128]).cuda().eval()
frames = torch.randint(0, 256, (3, 40, 56, 3), dtype=torch.uint8).cuda()
_, _, m = net.predict_mask(frames, size=(32, 48), want=("mask",))                      # tile kernel, span mode
same = torch.randint(0, 256, (2, 32, 48, 3), dtype=torch.uint8).cuda()
net.predict_mask(same, size=(32, 48), want=("mask",))                                   # same-size copy path
wide = torch.randint(0, 256, (1, 300, 401, 3), dtype=torch.uint8).cuda()
U.preprocess_u8(wide, size=(32, 48))                                                    # tile kernel, sparse mode, odd pitch
net.b200_precision = "fp32"
net.predict_mask(frames, size=(32, 48), want=("logits", "mask"))                        # fp32-class plan + fp32 preprocess
net.b200_precision = "bf16"
net3 = U.UNet(3, 3, [64, 128]).cuda().eval()
net3(torch.randn(2, 3, 32, 48).cuda())                                                  # multi-channel head
net.train()
step = U.FusedTrainStep(net, cuda_graph=False)
x = torch.randn(4, 3, 32, 32).cuda()
y = (torch.rand(4, 1, 32, 32) < 0.2).float().cuda()
print(step.step(x, y).tolist())
eng = step._last_eng                                                                    # staged backward, stage by stage
flat = net._b200_flat
g = torch.empty_like(flat)
dz = torch.randn(4, 32, 32, device="cuda")
from unet_lane_detection_b200.training import train_forward  # noqa: E402
x4 = U.ops.nchw_to_nhwc4(x)
eng2, _ = train_forward(net, x4)
st = torch.cuda.current_stream().cuda_stream
for s in range(eng2.n_stages):
    check(lib.unet_b200_train_backward_stage(eng2.handle, s, dz.data_ptr(), flat.data_ptr(), g.data_ptr(), st))
check(lib.unet_b200_trainer_join(eng2.handle, st, st))
small = U.UNet(3, 1, [32, 64]).cuda().train()                                           # zero-extended widths in the trainer
step2 = U.FusedTrainStep(small, cuda_graph=False)
x2 = torch.randn(8, 3, 32, 32).cuda()
y2 = (torch.rand(8, 1, 32, 32) < 0.2).float().cuda()
print(step2.step(x2, y2).tolist())
M = np.array([[1.1, 0.2, -3.0], [0.05, 0.9, 2.0], [0.0, -0.004, 1.0]])
x4w = U.ops.preprocess_warp_u8(frames, M, (70, 50), (32, 48))
up = U.ops.resize_gray_u8(m, (50, 70))
torch.cuda.synchronize()
print("ok", m.shape, x4w.shape, up.shape, float(g.abs().sum()))
