"""Run one layer configuration a few times (for ncu / timing). Usage:
   python tools/run_layer.py conv B H W C0 C1 Cout pool [reps] [halo]
   python tools/run_layer.py stem B H W [reps]
   python tools/run_layer.py convT B H W Cin f [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import unet_lane_detection_b200 as U
from unet_lane_detection_b200._lib import check, lib

kind = sys.argv[1]
dev = torch.device("cuda")
if kind == "conv":
    B, H, W, C0, C1, Cout, pool = [int(v) for v in sys.argv[2:9]]
    reps = int(sys.argv[9]) if len(sys.argv) > 9 else 5
    halo = int(sys.argv[10]) if len(sys.argv) > 10 else 1
    check(lib.unet_b200_set_option(b"halo", halo))
    x0 = torch.randn(B, H, W, C0, device=dev).to(torch.bfloat16)
    x1 = torch.randn(B, H, W, C1, device=dev).to(torch.bfloat16) if C1 else None
    w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) / (3.0 * (C0 + C1) ** 0.5)
    wp, bias = U.pack_conv3x3(w)
    fn = lambda: U.conv3x3(x0, wp, bias, x1=x1, pool=bool(pool))
    fl = 2.0 * B * H * W * Cout * 9 * (C0 + C1)
    nbytes = B * H * W * (C0 + C1 + Cout * (1.25 if pool else 1)) * 2
elif kind == "stem":
    B, H, W = [int(v) for v in sys.argv[2:5]]
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    x4 = torch.randn(B, H, W, 4, device=dev).to(torch.bfloat16)
    w = torch.randn(64, 3, 3, 3, device=dev) / 5
    wp, bias = U.pack_stem_tc(w)
    fn = lambda: U.stem_conv_tc(x4, wp, bias)
    fl = 2.0 * B * H * W * 64 * 27
    nbytes = B * H * W * (8 + 128)
else:
    B, H, W, Cin, f = [int(v) for v in sys.argv[2:7]]
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cin, f, 2, 2, device=dev) / Cin ** 0.5
    wp = U.pack_convT2x2(w)
    bias = torch.zeros(f, device=dev)
    fn = lambda: U.convT2x2(x, wp, bias)
    fl = 2.0 * B * H * W * 4 * f * Cin
    nbytes = B * H * W * (Cin + 4 * f) * 2
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{' '.join(sys.argv[1:])}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  {nbytes/ms/1e6:.0f} GB/s (algorithmic bytes)")
