"""Run one conv3x3 layer configuration a few times (for ncu / timing). Usage:
   python tools/run_layer.py B H W C0 C1 Cout pool [reps] [halo]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import unet_lane_detection_b200 as U
from unet_lane_detection_b200._lib import check, lib

B, H, W, C0, C1, Cout, pool = [int(v) for v in sys.argv[1:8]]
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
halo = int(sys.argv[9]) if len(sys.argv) > 9 else 1
check(lib.unet_b200_set_option(b"halo", halo))
dev = torch.device("cuda")
x0 = torch.randn(B, H, W, C0, device=dev).to(torch.bfloat16)
x1 = torch.randn(B, H, W, C1, device=dev).to(torch.bfloat16) if C1 else None
w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) / (3.0 * (C0 + C1) ** 0.5)
wp, bias = U.pack_conv3x3(w)
for _ in range(2):
    U.conv3x3(x0, wp, bias, x1=x1, pool=bool(pool))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    U.conv3x3(x0, wp, bias, x1=x1, pool=bool(pool))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * B * H * W * Cout * 9 * (C0 + C1)
print(f"conv B{B} {H}x{W} {C0}+{C1}->{Cout} pool={pool} halo={halo}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
