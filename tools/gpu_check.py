"""First-light checks on a B200 (run via gpurun). Each case runs in its own process so that a trap
in one kernel cannot poison the CUDA context of the others.

    python tools/gpu_check.py            # run all cases, each under `timeout`
    python tools/gpu_check.py <case>     # run one case in-process
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["stem", "conv_64_64_s32", "conv_shapes", "conv_dual", "conv_pool", "convT", "head", "preprocess", "pool",
         "halo_10_0", "halo_10_1", "halo_16_0", "halo_16_1", "unet_small", "unet_full"]


def stats(name, got, ref, tol):
    import torch
    got, ref = got.float().cpu(), ref.float().cpu()
    d = (got - ref).abs()
    rel = d.max().item() / (ref.abs().max().item() + 1e-12)
    ok = bool(d.max().item() <= tol * max(1.0, ref.abs().max().item()))
    print(f"  {name}: max|d|={d.max().item():.4e} mean|d|={d.mean().item():.3e} ref_max={ref.abs().max().item():.3e} "
          f"rel={rel:.3e} -> {'PASS' if ok else 'FAIL'}", flush=True)
    if not ok:
        idx = (d > tol * max(1.0, ref.abs().max().item())).nonzero()
        print(f"    {idx.shape[0]} bad of {d.numel()}; first bad idx {idx[:5].tolist()}", flush=True)
    return ok


def run_case(case):
    import numpy as np
    import torch
    import torch.nn.functional as F
    import unet_lane_detection_b200 as U
    from oracle import unet_oracle as O

    dev = torch.device("cuda")
    torch.manual_seed(0)
    bf = O.bf16_round
    ok = True

    def nhwc(x):  # NCHW fp32 cpu -> NHWC bf16 cuda
        return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)

    def nchw(y):  # NHWC cuda -> NCHW fp32 cpu
        return y.float().cpu().permute(0, 3, 1, 2).contiguous()

    def conv_case(B, H, W, C0, C1, Cout, pool=False, relu=True):
        x0 = bf(torch.randn(B, C0, H, W))
        x1 = bf(torch.randn(B, C1, H, W)) if C1 else None
        w = torch.randn(Cout, C0 + C1, 3, 3) / (3.0 * (C0 + C1) ** 0.5)
        g, b = torch.rand(Cout) + 0.5, torch.randn(Cout) * 0.1
        m, v = torch.randn(Cout) * 0.1, torch.rand(Cout) + 0.5
        wp, bias = U.pack_conv3x3(w.to(dev), (g.to(dev), b.to(dev), m.to(dev), v.to(dev), 1e-5))
        s = g / torch.sqrt(v + 1e-5)
        wf, bf_ = bf(w * s[:, None, None, None]), b - m * s
        xin = torch.cat([x0, x1], 1) if C1 else x0
        ref = F.conv2d(xin, wf, bf_, padding=1)
        if relu:
            ref = F.relu(ref)
        t0 = time.time()
        out = U.conv3x3(nhwc(x0), wp, bias, x1=nhwc(x1) if C1 else None, relu=relu, pool=pool)
        torch.cuda.synchronize()
        y = out[0] if pool else out
        r = stats(f"conv B{B} {H}x{W} {C0}+{C1}->{Cout} ({time.time()-t0:.2f}s)", nchw(y), bf(ref), 2e-2)
        if pool:
            r &= stats("  fused pool", nchw(out[1]), F.max_pool2d(nchw(y), 2), 0.0)
        return r

    if case == "stem":
        x = bf(torch.randn(2, 3, 32, 48))
        w = torch.randn(64, 3, 3, 3) / 5
        g, b, m, v = torch.rand(64) + 0.5, torch.randn(64) * 0.1, torch.randn(64) * 0.1, torch.rand(64) + 0.5
        ws, bias = U.pack_stem(w.to(dev), (g.to(dev), b.to(dev), m.to(dev), v.to(dev), 1e-5))
        s = g / torch.sqrt(v + 1e-5)
        ref = F.relu(F.conv2d(x, bf(w * s[:, None, None, None]), b - m * s, padding=1))
        x4 = U.nchw_to_nhwc4(x.to(dev))
        y = U.stem_conv(x4, ws, bias, 3)
        ok &= stats("stem 3->64", nchw(y), bf(ref), 1e-2)
    elif case == "conv_64_64_s32":
        ok &= conv_case(1, 32, 32, 64, 0, 64)
    elif case == "conv_shapes":
        for (B, H, C0, Co) in [(1, 224, 64, 64), (1, 112, 64, 128), (2, 56, 128, 256), (2, 56, 256, 256),
                               (8, 28, 256, 512), (32, 14, 512, 1024), (3, 14, 1024, 1024), (1, 16, 128, 128)]:
            ok &= conv_case(B, H, H, C0, 0, Co)
    elif case == "conv_dual":
        ok &= conv_case(1, 112, 112, 128, 128, 128)
        ok &= conv_case(2, 28, 28, 512, 512, 512)
        ok &= conv_case(1, 48, 80, 64, 64, 64)
    elif case == "conv_pool":
        ok &= conv_case(1, 224, 224, 64, 0, 64, pool=True)
        ok &= conv_case(2, 56, 56, 256, 0, 256, pool=True)
        ok &= conv_case(8, 28, 28, 512, 0, 512, pool=True)
        ok &= conv_case(1, 60, 80, 64, 0, 64, pool=True)
    elif case == "convT":
        for (B, H, Cin, f) in [(32, 14, 1024, 512), (2, 28, 512, 256), (1, 56, 256, 128), (1, 112, 128, 64)]:
            x = bf(torch.randn(B, Cin, H, H))
            w = torch.randn(Cin, f, 2, 2) / Cin ** 0.5
            b = torch.randn(f) * 0.1
            wp = U.pack_convT2x2(w.to(dev))
            y = U.convT2x2(nhwc(x), wp, b.to(dev))
            ref = F.conv_transpose2d(x, bf(w), b, stride=2)
            ok &= stats(f"convT B{B} {H}x{H} {Cin}->{f}", nchw(y), bf(ref), 2e-2)
    elif case == "head":
        x = bf(torch.randn(2, 64, 40, 56))
        w, b = torch.randn(64) / 8, 0.1
        lg, pr, mk = U.head(nhwc(x), w.to(dev), b, 0.5)
        ref = (x * w[None, :, None, None]).sum(1) + b
        ok &= stats("head logits", lg, ref, 1e-4)
        ok &= stats("head probs", pr, torch.sigmoid(ref), 1e-5)
        agree = ((mk.cpu() > 0) == (ref > 0)).float().mean().item()
        print(f"  mask agreement {agree:.6f}")
        ok &= agree > 0.9999
    elif case == "preprocess":
        import cv2
        rng = np.random.default_rng(0)
        for (hs, ws_) in [(480, 640), (224, 224), (685, 1055), (960, 1280)]:
            img = rng.integers(0, 256, (2, hs, ws_, 3), dtype=np.uint8)
            y, r = U.preprocess_u8(torch.from_numpy(img).to(dev), swap_rb=True, return_resized=True)
            ref_r = np.stack([cv2.resize(im, (224, 224))[:, :, ::-1] for im in img])
            exact = np.array_equal(r.cpu().numpy(), ref_r)
            refn = (ref_r.astype(np.float32) - np.array(U.ops.MEAN_255, np.float32)) / np.array(U.ops.STD_255, np.float32)
            dn = (y[..., :3].float().cpu() - bf(torch.from_numpy(refn))).abs().max().item()
            print(f"  preprocess {hs}x{ws_}: resized bit-exact={exact} normalised max|d|={dn:.3e} pad={y[..., 3].abs().max().item()}")
            ok &= exact and dn < 2e-2
    elif case == "pool":
        x = bf(torch.randn(2, 64, 28, 36))
        ok &= stats("maxpool", nchw(U.maxpool2x2(nhwc(x))), F.max_pool2d(x, 2), 0.0)
    elif case.startswith("halo_"):
        _, pitch, bo = case.split("_")
        lib = C.CDLL(os.path.join(ROOT, "tools", "libhalo_probe.so"))
        H, W = 48, 40
        x = bf(torch.randn(1, 64, H, W))
        w = torch.randn(64, 64, 3, 3) / 24
        wp, _ = U.pack_conv3x3(w.to(dev))
        xd = nhwc(x)
        ref = F.conv2d(x, bf(w), padding=1)
        for (h0, w0) in [(16, 8), (0, 0), (32, 32)]:
            out = torch.zeros(128, 64, device=dev)
            rc = lib.halo_probe(C.c_void_p(xd.data_ptr()), H, W, C.c_void_p(wp.data_ptr()), h0, w0, int(pitch), int(bo),
                                C.c_void_p(out.data_ptr()))
            if rc != 0:
                print(f"  halo_probe rc={rc}")
                ok = False
                break
            got = out.cpu().reshape(16, 8, 64).permute(2, 0, 1)  # [co][h][w]
            ok &= stats(f"halo pitch={pitch} bo={bo} tile@({h0},{w0})", got, ref[0, :, h0:h0 + 16, w0:w0 + 8], 2e-3)
    elif case in ("unet_small", "unet_full"):
        feats, hw, B = ([64, 128], (32, 48), 2) if case == "unet_small" else ([64, 128, 256, 512], (224, 224), 2)
        torch.manual_seed(0)
        ref = O.UNetOracle(3, 1, feats).eval()
        O.randomize_bn_(ref, 1)
        O.scale_head_(ref, 40.0)
        x = torch.randn(B, 3, *hw, generator=torch.Generator().manual_seed(1234))
        with torch.no_grad():
            y32 = ref(x)
            yem, feats_em = O.forward_bf16_emulated(ref, x, return_feats=True)
        net = U.UNet(3, 1, feats)
        net.load_state_dict(ref.state_dict())
        net = net.to(dev).eval()
        t0 = time.time()
        with torch.no_grad():
            y = net(x.to(dev))
        torch.cuda.synchronize()
        print(f"  forward took {time.time()-t0:.2f}s (first call, includes packing)")
        ok &= stats("logits vs bf16-emulated oracle", y, yem, 2e-2)
        ok &= stats("logits vs fp32 oracle", y, y32, 2e-2)
        print(f"  logit std {y32.std().item():.3f}; mask agreement vs fp32 oracle: {O.mask_agreement(y.cpu(), y32):.5f} "
              f"(excluding |z|<0.05: {O.mask_agreement(y.cpu(), y32, band=0.05):.5f}); "
              f"emulated-vs-fp32 floor: {O.mask_agreement(yem, y32):.5f}")
    else:
        raise SystemExit(f"unknown case {case}")
    print(f"CASE {case}: {'PASS' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1:
        sys.exit(run_case(sys.argv[1]))
    summary = []
    for c in CASES:
        print(f"=== {c}", flush=True)
        try:
            r = subprocess.run(["timeout", "240", sys.executable, os.path.abspath(__file__), c], cwd=ROOT)
            summary.append((c, r.returncode))
        except Exception as e:  # noqa: BLE001
            summary.append((c, repr(e)))
    print("SUMMARY", summary)


if __name__ == "__main__":
    main()
