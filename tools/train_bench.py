"""Time the fused training step (config 4: batch 64/GPU, 224x224, BCE+Dice, AdamW) and print a per-kernel table.
    python tools/train_bench.py [--batch 64] [--steps 10] [--profile]
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_lane_detection_b200 as U  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--hw", nargs=2, type=int, default=[224, 224])
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--per-launch", default=None, help="substring: list every launch of matching kernels in order")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="one AdamW launch after the whole backward")
    ap.add_argument("--timeline", default=None, help="write every kernel of the profiled step (start, duration, stream) to this JSON")
    ap.add_argument("--opt", action="append", default=[], help="library switch name=value (unet_b200_set_option)")
    args = ap.parse_args()
    from unet_lane_detection_b200._lib import check, lib
    for o in args.opt:
        k, v = o.split("=")
        check(lib.unet_b200_set_option(k.encode(), int(v)))
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128, 256, 512]).cuda().train()
    B, (H, W) = args.batch, args.hw
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn(B, 3, H, W, device="cuda", generator=g)
    y = (torch.rand(B, 1, H, W, device="cuda", generator=g) < 0.085).float()
    step = U.FusedTrainStep(net, cuda_graph=not args.eager, overlap=not args.no_overlap)
    for _ in range(3):
        losses = step.step(x, y)
    torch.cuda.synchronize()
    print("warm losses", losses.tolist())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = step.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = 3 * 73.756e9 * B * (H * W) / (224 * 224)
    print(f"train step: {ms:.3f} ms/step, {B / ms * 1e3:.1f} samples/s, {flops / ms / 1e9:.1f} TFLOP/s (3x fwd FLOPs), losses {losses.tolist()}")
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step.step(x, y)
            torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                k = ev.name[:90]
                t, n = agg.get(k, (0.0, 0))
                agg[k] = (t + ev.device_time_total if hasattr(ev, "device_time_total") else t + ev.cuda_time_total, n + 1)
        if args.timeline:
            import json
            rows = []
            for ev in prof.profiler.kineto_results.events():
                if ev.device_type() == torch.autograd.DeviceType.CUDA:
                    rows.append({"name": ev.name()[:60], "start_us": ev.start_ns() / 1e3, "dur_us": ev.duration_ns() / 1e3,
                                 "stream": ev.device_resource_id()})
            rows.sort(key=lambda r: r["start_us"])
            t0 = rows[0]["start_us"] if rows else 0.0
            for r in rows:
                r["start_us"] -= t0
            with open(args.timeline, "w") as f:
                json.dump(rows, f)
        if args.per_launch:
            evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and args.per_launch in ev.name]
            evs.sort(key=lambda e: e.time_range.start)
            for i, ev in enumerate(evs):
                print(f"  #{i:2d} {ev.device_time_total:8.1f} us  {ev.name[:70]}")
        tot = sum(t for t, _ in agg.values())
        for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}% x{n:3d}  {k}")
        print(f"total kernel time {tot / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
