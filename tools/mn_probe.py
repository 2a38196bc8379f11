import ctypes as C, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for v in (0, 1):
        subprocess.run(["timeout", "60", sys.executable, __file__, str(v)])
    sys.exit(0)
import torch
v = int(sys.argv[1])
lib = C.CDLL(os.path.join(ROOT, "tools", "libmn_probe.so"))
torch.manual_seed(0)
A = torch.randn(128, 128).to(torch.bfloat16).cuda()
B = torch.randn(128, 64).to(torch.bfloat16).cuda()
out = torch.zeros(128, 64, device="cuda")
rc = lib.mn_probe(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), v, C.c_void_p(out.data_ptr()))
ref = A.float().t() @ B.float()
print(f"variant {v}: rc={rc} max|d|={(out - ref).abs().max().item():.4e} ref_max={ref.abs().max().item():.2f}")
