"""Per-kernel device times of ONE frame through the plan (the executor's per-frame path), best of 20 profiled passes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
net = U.UNet(3, 1, [64, 128, 256, 512]).cuda().eval()
x4 = torch.randn(B, 224, 224, 4, device="cuda").to(torch.bfloat16)
best = None
for _ in range(20):
    rows = net.profile_layers(x4)
    if best is None:
        best = rows
    else:
        for b, r in zip(best, rows):
            b["ms"] = min(b["ms"], r["ms"])
tot = sum(r["ms"] for r in best)
print(f"batch {B}: sum of kernel times {tot * 1e3:.1f} us")
for r in best:
    fl = ("P" if r["fused_pool"] else "-") + ("H" if r["halo"] else "-") + ("F" if r["fused_head"] else "-")
    tf = r["flops"] / r["ms"] / 1e9 if r["ms"] > 0 else 0
    print(f"{r['kind']:9s} {r['H']:4d}x{r['W']:<4d} {r['Cin']:5d}->{r['Cout']:<5d} bn={r['block_n']:<4d} {fl} {r['ms'] * 1e3:7.1f} us {tf:7.1f} TF/s")
