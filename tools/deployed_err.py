"""Spread of the eval-vs-oracle error after 31 training steps of the deployed topology (test_deployed_topology_trains),
with the dgrad-epilogue BatchNorm-backward sums on and off."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_lane_detection_b200 as U  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402
from unet_lane_detection_b200._lib import check, lib  # noqa: E402

for fuse in (0, 1, 0, 1, 0, 1):
    check(lib.unet_b200_set_option(b"dgrad_fuse", fuse))
    torch.manual_seed(0)
    ref = O.UNetOracle(3, 1, [32, 64, 128]).train()
    O.randomize_bn_(ref, seed=1)
    net = U.UNet(3, 1, [32, 64, 128])
    net.load_state_dict(ref.state_dict())
    net = net.cuda().train()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 64, 64, generator=g).cuda()
    y = torch.zeros(8, 1, 64, 64)
    y[:, :, 16:48, 24:40] = 1.0
    y = y.cuda()
    step = U.FusedTrainStep(net, lr=1e-3)
    first = step.step(x, y).cpu()
    for _ in range(30):
        last = step.step(x, y)
    last = last.cpu()
    net.eval()
    ref.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    ref.eval()
    with torch.no_grad():
        got, want = net(x).cpu(), ref(x.cpu())
        emu = O.forward_bf16_emulated(ref, x.cpu())
    print(f"fuse={fuse} first={first[0].item():.4f} last={last[0].item():.4f} err={(got - want).abs().max().item():.4f} "
          f"emulated_err={(emu - want).abs().max().item():.4f} range={want.abs().max().item():.3f}")
