"""Write-rate ceiling of this GPU, to judge the store-bound kernels (stem, ConvT at 112->224) against:
plain fills (cudaMemset through torch.zero_, an elementwise fill kernel) and a copy, 2 GiB each, best of 5, CUDA events.
    python tools/fill_probe.py            # prints one JSON line"""
import json

import torch

n = 1 << 30   # bf16 elements: 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")


def best(fn, reps=5):
    t = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t


out = {}
for name, fn, bytes_ in (("memset_zero", lambda: a.zero_(), 2 * n), ("fill_kernel", lambda: a.fill_(1.5), 2 * n),
                         ("copy_read_plus_write", lambda: b.copy_(a), 4 * n), ("read_only_sum", lambda: a.view(torch.int16).max(), 2 * n)):
    fn()
    ms = best(fn)
    out[name] = {"ms": ms, "gbs": bytes_ / (ms / 1e3) / 1e9}
print(json.dumps(out))
