#!/usr/bin/env python
"""Headline benchmark: U-Net 224x224 inference frames/s on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--chunk C]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's own CPU PyTorch pipeline (oracle)

A step = one pass of the hot path over one batch of synthetic frames on every rank:
uint8 frames -> fused resize/normalise -> U-Net (folded BN, bf16 tensor-core convs) -> sigmoid ->
threshold -> uint8 mask.  `value` times that with the frames already in HBM; `e2e` times the
reference-facing host-buffer call (pinned host frames in, host masks out, copies inside the timed
region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import numpy as np

# One process per GPU, every GPU visible to every rank: the library keeps its state per device and the NVLink / NVSwitch
# gradient exchange of the training step maps the peers' buffers, so nothing is pinned by default. UB_BENCH_PIN=1 restricts
# each rank to its own device before CUDA initialises (the training step then falls back to NCCL).
PINNED = "LOCAL_RANK" in os.environ and os.environ.get("UB_BENCH_PIN", "0") == "1"
if PINNED:
    _vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    _ids = _vis.split(",") if _vis else None
    _lr = int(os.environ["LOCAL_RANK"])
    os.environ["CUDA_VISIBLE_DEVICES"] = _ids[_lr] if _ids and _lr < len(_ids) else str(_lr)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unet224_inference_frames_per_sec"
UNIT = "frames/s"
FEATURES = [64, 128, 256, 512]
FLOPS_PER_FRAME = 73.756e9  # conv3x3 + convT + 1x1, 2*MAC, 224x224, features [64..512]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": p.get("bf16_tflops_sustained", 1400.0), "bf16_burst": p.get("bf16_tflops", 1590.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def build_model_cpu(seed=0):
    """Random-init reference architecture with non-trivial BatchNorm statistics."""
    import torch
    import unet_lane_detection_b200 as U
    torch.manual_seed(seed)
    m = U.UNet(3, 1, FEATURES)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                n = mod.num_features
                mod.weight.copy_(torch.rand(n, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(n, generator=g) * 0.1)
                mod.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(n, generator=g) + 0.5)
    return m.eval()


def physical_gpu_index(ordinal=0):
    """Index nvidia-smi / NVML know this process's GPU `ordinal` by (CUDA_VISIBLE_DEVICES may hold indices or UUIDs)."""
    vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    first = vis[ordinal] if ordinal < len(vis) else str(ordinal)
    try:
        return int(first)
    except ValueError:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=index,uuid", "--format=csv,noheader"], capture_output=True, text=True,
                                 timeout=10).stdout
            for line in out.splitlines():
                idx, uuid = [c.strip() for c in line.split(",")]
                if uuid.startswith(first) or first.startswith(uuid):
                    return int(idx)
        except Exception:  # noqa: BLE001
            pass
        return 0


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled WHILE the timed region runs: NVML in-process every 20 ms (the GPU index is the
    physical one, NVML ignores CUDA_VISIBLE_DEVICES); `nvidia-smi` polling as the fallback when pynvml is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.sm_max, self.power = None, []
        self.source = "nvidia-smi"

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        while not self._halt.is_set():
            mask = int(reasons_fn(h))
            self.rows.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(self.sm_max), "",
                              *["Active" if mask & self.BITS[n] else "Not Active" for n in self.NAMES]])
            try:
                self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.02)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:  # noqa: BLE001  (no pynvml / NVML error: poll nvidia-smi instead)
            pass
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def summary(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for n, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
               "reasons": sorted(reasons), "samples": len(self.rows), "source": self.source}
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


def cpu_reference_pipeline(model_cpu, frames_u8_nhwc, threshold=0.5):
    """The reference's CPU path (src/unet.py:24-72 around README.md:1460-1481), restated in oracle/."""
    import torch
    from oracle import unet_oracle as O
    pre = [O.preprocess_oracle(f, (224, 224), swap_rb=True)[0] for f in frames_u8_nhwc]   # BGR->RGB + resize (identity at 224) + batch dim
    import numpy as np
    x = torch.from_numpy(O.normalize_oracle(np.concatenate(pre, 0)))
    with torch.no_grad():
        logits = model_cpu(x).numpy()
    return [O.postprocess_oracle([logits[i:i + 1]], (224, 224), threshold) for i in range(len(frames_u8_nhwc))]


def time_cpu_reference(frames_per_step, steps, warmup, threads=None):
    """frames/s of the oracle pipeline on the host cores, all threads (or `threads`)."""
    import numpy as np
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.manual_seed(0)
    m = O.UNetOracle(3, 1, FEATURES).eval()
    O.randomize_bn_(m, 1)
    rng = np.random.default_rng(1234)
    frames = rng.integers(0, 256, (frames_per_step, 224, 224, 3), dtype=np.uint8)
    for _ in range(warmup):
        cpu_reference_pipeline(m, frames)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pipeline(m, frames)
    dt = time.perf_counter() - t0
    return frames_per_step * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps_frames = 8
    value, sec_per_step, threads = time_cpu_reference(fps_frames, args.steps, args.warmup)
    sample = f"{fps_frames} frames/step x {args.steps} steps, oracle fp32 PyTorch CPU pipeline (resize+normalise+UNet+sigmoid+threshold)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "U-Net 224x224 inference, features [64,128,256,512], synthetic uint8 frames; bounded CPU sample",
                   "frames_per_step": fps_frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


TRAIN_FLOPS_PER_SAMPLE = 3 * FLOPS_PER_FRAME  # forward + dgrad + wgrad


def measure_train(dev, world, rank, batch, steps, warmup, timed, host_inputs=False, exchange="auto", overlap=True):
    """BASELINE.json configs[3]: U-Net training step (BCE+Dice, AdamW), bf16 tensor-core convs, `batch` samples per GPU,
    NCCL all-reduce of the flat fp32 gradient when world > 1. Returns a dict for the JSON line."""
    import torch
    import unet_lane_detection_b200 as U
    torch.manual_seed(0)
    net = U.UNet(3, 1, FEATURES).to(dev).train()
    g = torch.Generator(device="cpu").manual_seed(42 + rank)           # README.md:2248 seed + rank
    x_host = torch.randn(batch, 3, 224, 224, generator=g).pin_memory()
    y_host = (torch.rand(batch, 1, 224, 224, generator=g) < 0.085).float().pin_memory()   # 8.5 % positives (README.md:2534)
    x, y = x_host.to(dev), y_host.to(dev)
    step = U.FusedTrainStep(net, exchange=exchange, overlap=overlap)    # lr 1e-4, wd 1e-4, pos_weight 3 (README.md:2169-2174)
    box = {}
    check = None
    if world > 1:
        # proof that the exchange ran and is right, taken on the first (eager) step: the gradient sum exactly as the exchange
        # kernel forms it (NVLink peer loads / NVSwitch multimem.ld_reduce / the bucketed NCCL all-reduce) against one plain
        # all-reduce of copies of the local gradients, and the parameters of all replicas against each other afterwards
        import torch.distributed as dist
        step.step(x, y)
        torch.cuda.synchronize()
        if step.nvlink is not None and step.nvlink.mode != "push":
            ref = step.grads.clone()
            dist.all_reduce(ref)
            num = torch.zeros(1, device=dev)
            for a, b, t in step.nvlink.reduced_parts():
                num = torch.maximum(num, (t - ref[a:b]).abs().max().reshape(1))
            dist.all_reduce(num, op=dist.ReduceOp.MAX)
            check = float(num.item()) / float(ref.abs().max().item())
        elif step.nvlink is None:
            # NCCL buckets reduce in place: compare the bucketed result with an all-reduce of a second, un-bucketed backward
            # is not possible without re-running; report the spread of the reduced buffer across ranks instead (must be 0)
            lo, hi = step.grads.clone(), step.grads.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            check = float((hi - lo).abs().max().item()) / float(hi.abs().max().item())
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        pl, ph = flat.clone(), flat.clone()
        dist.all_reduce(pl, op=dist.ReduceOp.MIN)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        box["params_spread"] = float((ph - pl).abs().max().item())

    def run_dev():
        box["loss"] = step.step(x, y)

    def run_host():                                                     # README.md:2067-2081: .to(device) ... loss.item()
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        box["loss_host"] = step.step(xd, yd).tolist()

    for _ in range(max(warmup, 3)):
        run_dev()
    ms = timed(run_dev, steps)
    loss = box["loss"].tolist()
    value = world * batch * steps / (ms / 1e3)
    out = {"metric": "unet224_train_samples_per_sec", "value": value, "unit": "samples/s", "ms_per_step": ms / steps,
           "batch_per_gpu": batch, "tflops_per_gpu": value / world * TRAIN_FLOPS_PER_SAMPLE / 1e12,
           "flops_per_sample": TRAIN_FLOPS_PER_SAMPLE, "loss_after": loss,
           "exchange": step.exchange,
           "exchange_check": check,
           "exchange_check_note": None if world == 1 else (
               "step 1: max |sum formed by the exchange kernel - all_reduce(local gradients)| / max |gradient|, max over ranks"
               if step.nvlink is not None else "step 1: spread of the bucket-wise all-reduced gradient across ranks / max |gradient|"),
           "params_spread_after_step1": box.get("params_spread"),
           "buckets": None if world == 1 else [[int(s_), int(a), int(b)] for s_, a, b in (step.buckets or [])],
           "overlap": None if world == 1 else ("each bucket's exchange + AdamW runs on a side stream as soon as its backward stages have finished"
                                              if overlap else "off: one exchange after the whole backward (A/B run)"),
           "collective": "none" if world == 1 else (
               f"NCCL all-reduce(sum) per gradient bucket ({31037633 * 4 / 1e6:.0f} MB per step in total) + AdamW on the bucket" if step.exchange == "nccl" else
               "none: per gradient bucket ONE kernel sums its part over the peers' buffers (NVLink loads), applies AdamW (ZeRO-1 "
               "sharded state) and stores the parameters to all replicas (NVLink stores); device-side barriers"
               if step.exchange == "nvlink" else
               "none: per gradient bucket ONE kernel on NVSwitch multicast addresses - multimem.ld_reduce sums the gradients inside the "
               "switch, AdamW (ZeRO-1 sharded state), multimem.st broadcasts the parameters; device-side barriers"
               if step.exchange == "nvlink_mc" else
               "none: NVLink P2P gradient atomics to the owner replica inside the backward kernels, sharded AdamW, parameter "
               "stores to all replicas; two device-side barriers per step"),
           "config": "BASELINE.json configs[3]: BCE+Dice (pos_weight 3), AdamW lr 1e-4 wd 1e-4, BatchNorm batch statistics per replica"}
    if host_inputs:
        run_host()
        ms_h = timed(run_host, steps)
        out["e2e"] = {"value": world * batch * steps / (ms_h / 1e3), "unit": "samples/s", "ms_per_step": ms_h / steps,
                      "h2d_bytes_per_step": batch * 224 * 224 * 4 * 4, "d2h_bytes_per_step": 12,
                      "api": "FusedTrainStep.step on pinned host tensors, loss read back every step"}
    del step, net
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150, help="timed steps (default: a timed region of about 2 s at N = 1)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None,
                    help="frames per GPU per step. Default: 256 at N = 1 (BASELINE.json configs[1]); 4096 / N at N > 1 "
                         "(configs[2]: batch 4096 sharded across the GPUs, strong scaling)")
    ap.add_argument("--src-hw", nargs=2, type=int, default=None, metavar=("HS", "WS"),
                    help="source frame size when it differs from the network input (480 640 = camera frames, a real resize in "
                         "the fused preprocess); by default the headline uses frames of the network size and the 480x640 case "
                         "is reported as the extra key e2e_src480x640")
    ap.add_argument("--chunk", type=int, default=256, help="frames per pass through the plan (256 = the whole batch in one pass)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: BASELINE.json headline (configs[1]); train: the training step of configs[3] as the metric")
    ap.add_argument("--train-batch", type=int, default=64, help="samples per GPU per training step (configs[3])")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "nvlink", "nvlink_pull", "nvlink_mc", "nvlink_push"],
                    help="training, N > 1: gradient exchange (auto = nvlink when the ranks can map each other's memory). nvlink: one kernel = reduce-scatter by NVLink loads + sharded AdamW + "
                         "all-gather by NVLink stores; nvlink_push: gradient atomics go to the owner GPU inside the backward kernels")
    ap.add_argument("--no-train", action="store_true", help="infer mode: skip the secondary training-step measurement")
    ap.add_argument("--no-overlap", action="store_true", help="train mode, N > 1: one gradient exchange after the whole backward (A/B)")
    ap.add_argument("--hw", nargs=2, type=int, default=[224, 224], metavar=("H", "W"),
                    help="network input size; 480 640 = BASELINE.json configs[4] (camera resolution, use --batch 16 --chunk 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="infer mode: skip the fp32-class plan's throughput line")
    ap.add_argument("--no-cfg5", action="store_true", help="infer mode: skip the 480x640 (BASELINE.json configs[4]) secondary measurement")
    ap.add_argument("--layers-out", default=None, help="write the per-kernel profile table to this JSON file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU pipeline)")
    torch.cuda.set_device(0 if PINNED else int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = build_model_cpu().to(dev)
    model.b200_chunk = args.chunk
    strong = args.batch is None and world > 1 and args.mode == "infer"
    B = args.batch if args.batch is not None else (4096 // world if world > 1 else 256)
    H, W = args.hw
    Hs, Ws = args.src_hw if args.src_hw else (H, W)
    flops_per_frame = FLOPS_PER_FRAME * (H * W) / (224 * 224)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    frames_host = torch.randint(0, 256, (B, Hs, Ws, 3), dtype=torch.uint8, generator=g).pin_memory()
    frames_dev = frames_host.to(dev)
    mask_host = torch.empty(B, H, W, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if args.mode == "train":
        peaks = load_peaks()
        sampler = ClockSampler(index=physical_gpu_index(0 if PINNED else torch.cuda.current_device()) if rank == 0 else 0)
        if rank == 0:
            sampler.start()
        tr = measure_train(dev, world, rank, args.train_batch, args.steps, args.warmup, timed, host_inputs=True,
                           exchange=args.exchange, overlap=not args.no_overlap)
        clocks = sampler.summary() if rank == 0 else None
        line = {"metric": tr["metric"], "value": tr["value"], "unit": tr["unit"], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tr["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": tr["config"], "batch_per_gpu": args.train_batch, "parallelism": f"data-parallel x{world}",
                           "collective": tr["collective"],
                           "l2_policy": f"activations + gradients {args.train_batch * 3 * 64.1:.0f} MB per step >> 126 MB L2"},
                "e2e": tr["e2e"], "gpu_launches": None, "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": tr["tflops_per_gpu"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                             "frac": tr["tflops_per_gpu"] / peaks["bf16_sustained"], "traffic": None,
                             "kernel": "whole training step (forward + dgrad + wgrad GEMMs = 3x forward FLOPs; BN/loss/AdamW passes are HBM-bound extras)",
                             "peak_source": peaks["source"] + " bf16_tflops_sustained"},
                "loss_after": tr["loss_after"]}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            os.dup2(2, 1)
            dist.destroy_process_group()
        return

    def step_device():
        model.predict_mask(frames_dev, threshold=0.5, swap_rb=True, size=(H, W), want=("mask",))

    def step_host():
        model.infer_host(frames_host, threshold=0.5, swap_rb=True, size=(H, W), mask_out=mask_host)

    # pure-write ceiling of this GPU (a kernel that only writes cannot reach the copy bandwidth of MEASURED_PEAKS): a 2 GiB
    # cudaMemset, best of 5 after a warm-up, taken BEFORE the long loops (boost clocks, like the per-kernel event timings)
    fill_buf = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
    fill_buf.zero_()
    torch.cuda.synchronize()
    fill_ms = 1e30
    for _ in range(5):
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fe0.record()
        fill_buf.zero_()
        fe1.record()
        torch.cuda.synchronize()
        fill_ms = min(fill_ms, fe0.elapsed_time(fe1))
    write_peak = fill_buf.numel() * 2 / (fill_ms / 1e3) / 1e9
    del fill_buf

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(index=physical_gpu_index(0 if PINNED else torch.cuda.current_device()) if rank == 0 else 0)
    if rank == 0:
        sampler.start()
    l0 = model.gpu_launches
    ms = timed(step_device, args.steps)
    launches = model.gpu_launches - l0
    clocks = sampler.summary() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    for _ in range(2):
        step_host()
    ms_host = timed(step_host, args.steps)
    e2e_value = world * B * args.steps / (ms_host / 1e3)

    # second end-to-end figure: 480x640 camera frames (src/unet_ros_node.py publishes 640x480 bgr8, README.md:3763-3765), i.e.
    # a REAL bilinear resize in the fused preprocess and 6.1x the bytes over PCIe; the headline's source frames already have
    # the network size (configs[1] names 224x224 frames), so its resize is the identity
    e2e_cam = None
    if (Hs, Ws) == (H, W) == (224, 224):
        cam_host = torch.randint(0, 256, (B, 480, 640, 3), dtype=torch.uint8, generator=g).pin_memory()

        def step_cam():
            model.infer_host(cam_host, threshold=0.5, swap_rb=True, size=(H, W), mask_out=mask_host)

        for _ in range(2):
            step_cam()
        ms_cam = timed(step_cam, args.steps)
        e2e_cam = {"value": world * B * args.steps / (ms_cam / 1e3), "unit": UNIT, "ms_per_step": ms_cam / args.steps,
                   "h2d_bytes_per_step": B * 480 * 640 * 3, "d2h_bytes_per_step": B * H * W,
                   "source": "uint8 BGR 480x640 frames in pinned host memory -> cv2-exact bilinear resize to 224x224 + BGR->RGB + "
                             "normalise fused in the preprocess kernel"}
        del cam_host

    # ---- roofline ---------------------------------------------------------------------------------------------------------
    # Per-kernel CUDA events over one chunk-sized pass give every kernel's SHARE of the pass; the dominant family's in-step
    # time is that share x the driver-style timed ms_per_step (same clocks, same power state as `value`), and its fraction is
    # taken against the SUSTAINED bf16 peak. The per-kernel event times themselves come from short passes at boost clocks:
    # they are reported separately against the BURST peak (a kernel timed alone).
    peaks = load_peaks()
    nb = min(args.chunk, B)
    x4 = torch.empty(nb, H, W, 4, dtype=torch.bfloat16, device=dev).normal_()
    model.profile_layers(x4)
    rows = None
    for _ in range(3):
        r = model.profile_layers(x4)
        if rows is None:
            rows = r
        else:
            for a_, b_ in zip(rows, r):
                a_["ms"] = min(a_["ms"], b_["ms"])

    def time_kernel(fn, reps=3, inner=20):
        """ms per call: `inner` back-to-back calls between two events (a single short kernel would mostly measure the Python
        launch path: the GPU idles between the first event and the arrival of the launch), best of `reps`."""
        fn()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(inner):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / inner)
        return best

    import unet_lane_detection_b200 as U
    from unet_lane_detection_b200._lib import check as ub_check, f3 as ub_f3, lib as ub_lib
    from unet_lane_detection_b200.ops import MEAN_255, STD_255
    pre_out = torch.empty(nb, H, W, 4, dtype=torch.bfloat16, device=dev)

    def run_pre(src, hs, ws):       # the C-ABI call itself, output preallocated: the Python wrapper's allocation would dominate a 30 us kernel
        ub_check(ub_lib.unet_b200_preprocess_u8(src.data_ptr(), int(src.shape[0]), hs, ws, ws * 3, hs * ws * 3, H, W, 1, ub_f3(MEAN_255),
                                                ub_f3(STD_255), pre_out.data_ptr(), None, torch.cuda.current_stream().cuda_stream))

    pre_src = frames_dev[:nb].contiguous()
    pre_ms = time_kernel(lambda: run_pre(pre_src, Hs, Ws))
    all_ms = sum(r["ms"] for r in rows) + pre_ms
    passes_per_step = B / nb

    def fam(sel):
        rs = [r for r in rows if sel(r)]
        fl, ms_ = sum(r["flops"] for r in rs), sum(r["ms"] for r in rs)
        share = ms_ / all_ms
        in_step_ms = share * (ms / args.steps) / passes_per_step     # this family's time inside one pass of the timed region
        return {"launches": len(rs), "share_of_step": share, "flops": fl, "ms_burst": ms_,
                "tflops_burst": fl / (ms_ / 1e3) / 1e12 if ms_ > 0 else 0.0, "frac_of_burst_peak": fl / (ms_ / 1e3) / 1e12 / peaks["bf16_burst"] if ms_ > 0 else 0.0,
                "ms_in_step": in_step_ms, "tflops_in_step": fl / (in_step_ms / 1e3) / 1e12 if in_step_ms > 0 else 0.0,
                "frac_of_sustained_peak": fl / (in_step_ms / 1e3) / 1e12 / peaks["bf16_sustained"] if in_step_ms > 0 else 0.0}

    # dominant kernel = conv_umma2_kernel<256>, the CTA-pair implicit GEMM (over half of the step, see profiles/)
    dom = fam(lambda r: r["kind"] in ("conv3x3", "convT2x2") and r["block_n"] == 256 and not r.get("halo"))
    halo64 = fam(lambda r: r.get("halo") and r["block_n"] == 64)
    halo128 = fam(lambda r: r.get("halo") and r["block_n"] == 128)
    allconv = fam(lambda r: r["kind"] in ("conv3x3", "convT2x2"))
    # DRAM bytes of the dominant kernel family from the committed ncu capture (bench.py cannot run under ncu itself):
    # sum over its launches of one pass, i.e. the same unit as flops_per_pass; only valid for the captured geometry
    traffic, traffic_src = None, None
    for tname in ("r2_dram_traffic.json", "r1_dram_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and (H, W) == (224, 224):
            with open(tpath) as f:
                tj = json.load(f)
            fam_t = tj["families"].get("conv_umma2_kernel<256>")
            if fam_t and fam_t["launches"] == dom["launches"] and tj.get("chunk") == nb:
                traffic = fam_t["dram_read_bytes"] + fam_t["dram_write_bytes"]
                traffic_src = f"profiles/{tname} (ncu dram__bytes_read.sum + dram__bytes_write.sum, summed over the family's launches of one pass)"
                break

    # bandwidth-bound kernels: algorithmic bytes / CUDA-event time against the measured copy bandwidth (MEASURED_PEAKS hbm_gbs:
    # a copy, half reads half writes) and, for the write-dominated ones, their write rate against this GPU's pure-write
    # ceiling measured here (a 2 GiB cudaMemset, best of 3: ~3.9 TB/s on B200 - a kernel that only writes cannot reach the copy figure)
    def hbm_row(name, rd, wr, ms_, note):
        gbs = (rd + wr) / (ms_ / 1e3) / 1e9 if ms_ > 0 else 0.0
        wgbs = wr / (ms_ / 1e3) / 1e9 if ms_ > 0 else 0.0
        return {"kernel": name, "bytes": rd + wr, "read_bytes": rd, "write_bytes": wr, "ms": ms_, "gbs": gbs, "frac": gbs / peaks["hbm_gbs"],
                "write_gbs": wgbs, "frac_of_write_ceiling": wgbs / write_peak, "note": note}

    hbm = [hbm_row("preprocess_copy_u8_kernel" if (Hs, Ws) == (H, W) else "preprocess_u8_kernel", nb * 3 * Hs * Ws, nb * H * W * 8, pre_ms,
                   f"{nb} frames: reads 3*{Hs}*{Ws} B uint8, writes {H}*{W}*8 B NHWC4 bf16 per frame "
                   f"(algorithmic 3-channel output would be {H * W * 6} B)")]
    if (Hs, Ws) == (H, W) == (224, 224):
        ncam = nb
        cam_dev = torch.randint(0, 256, (ncam, 480, 640, 3), dtype=torch.uint8, device=dev)
        cam_ms = time_kernel(lambda: run_pre(cam_dev, 480, 640))
        # down-scaling by more than 2 skips source rows: 448 of a camera frame's 480 rows hold a tap of the 224 output rows
        f_ = ((np.arange(H) + 0.5) * (480 / H) - 0.5).astype(np.float32)
        s0_ = np.floor(f_).astype(np.int64)
        rows_t = len(set(np.clip(s0_, 0, 479)) | set(np.clip(s0_ + 1, 0, 479)))
        hbm.append(hbm_row("preprocess_bulk_u8_kernel (480x640 -> 224x224, cv2-exact bilinear)", ncam * 3 * rows_t * 640, ncam * H * W * 8, cam_ms,
                           f"{ncam} camera frames: the real-resize case of e2e_src480x640; read bytes = the {rows_t} of 480 source rows "
                           f"per frame that hold a tap (whole frames would be {ncam * 3 * 480 * 640} B)"))
        del cam_dev
    for r in rows:
        if r["kind"] == "stem":
            hbm.append(hbm_row("stem_umma_kernel", nb * r["H"] * r["W"] * 8, nb * r["H"] * r["W"] * r["Cout"] * 2, r["ms"],
                               "reads NHWC4 input, writes 64-channel bf16 output"))
        if r["kind"] == "convT2x2" and r["H"] >= H // 4:
            hbm.append(hbm_row(f"conv_umma2_kernel<256> ConvT {r['H']}x{r['W']} {r['Cin']}->{r['Cout']}",
                               nb * r["H"] * r["W"] * r["Cin"] * 2, nb * r["H"] * r["W"] * 4 * r["Cout"] * 2, r["ms"],
                               "reads the low-resolution tensor, writes the 2x up-sampled one (4 strided quad views)"))
    roofline = {"bound": "tensor", "achieved": dom["tflops_in_step"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": dom["frac_of_sustained_peak"], "traffic": traffic, "traffic_unit": "bytes per pass", "traffic_source": traffic_src,
                "kernel": f"conv_umma2_kernel<256> (cta_group::2; {dom['launches']} launches per pass: 3x3 convs with Cout>=256 + 4 ConvT)",
                "peak_source": peaks["source"] + " bf16_tflops_sustained",
                "method": "achieved = the family's algorithmic FLOPs per pass / (its share of a pass x the timed ms_per_step / passes per step); "
                          "share from CUDA events around every kernel of one pass on the launching stream (min of 3)",
                "burst": {"achieved": dom["tflops_burst"], "peak": peaks["bf16_burst"], "frac": dom["frac_of_burst_peak"],
                          "method": "the same FLOPs / the summed per-kernel CUDA-event times of a short pass (boost clocks) against bf16_tflops (burst)"},
                "share_of_step": dom["share_of_step"], "flops_per_pass": dom["flops"], "ms_per_pass_in_step": dom["ms_in_step"],
                "ms_per_pass_burst": dom["ms_burst"], "frames_per_pass": nb,
                "other_kernels": {"conv_halo2_kernel<64>": halo64, "conv_halo2_kernel<128>": halo128, "all_tensor_core_convs": allconv},
                "hbm": hbm, "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_write_ceiling_gbs": write_peak,
                "hbm_write_ceiling_how": "2 GiB cudaMemset on this GPU at the start of this run, best of 5, CUDA events",
                "whole_net_frac_of_peak": (value / world) * flops_per_frame / 1e12 / peaks["bf16_sustained"]}
    if args.layers_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
        with open(args.layers_out, "w") as f:
            json.dump({"chunk": int(x4.shape[0]), "rows": rows, "peaks": peaks, "preprocess_ms": pre_ms}, f, indent=1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"U-Net {H}x{W} bf16 inference, features {FEATURES}, "
                               + (f"batch {B * world} sharded across {world} GPUs ({B}/GPU), " if strong else f"batch {B}/GPU, ")
                               + f"uint8 {Hs}x{Ws} source frames, fused preprocess + mask threshold "
                               + (("(BASELINE.json configs[2])" if strong else "(BASELINE.json configs[1])") if (H, W) == (224, 224)
                                  else "(BASELINE.json configs[4] geometry)"),
                   "batch_per_gpu": B, "global_batch": B * world, "chunk": args.chunk, "source_hw": [Hs, Ws],
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2_policy": f"inputs+activations >> L2: {B * Hs * Ws * 3 / 1e6:.0f} MB frames and "
                                f"{min(args.chunk, B) * 64.1 * H * W / (224 * 224):.0f} MB activations per chunk vs 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * Hs * Ws * 3, "d2h_bytes_per_step": B * H * W,
                "ms_per_step": ms_host / args.steps, "api": "UNet.infer_host -> unet_b200_infer_u8_host_stream (pinned host buffers; input copies, preprocess and the first two layers pipelined piece by piece at the front of a pass, the fused-head conv and the mask copies piece by piece at its end)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if e2e_cam is not None:
        line["e2e_src480x640"] = e2e_cam
    if (H, W) != (224, 224):
        line["metric"] = f"unet{H}x{W}_inference_frames_per_sec"
        args.no_train = True
    elif not args.no_fp32:
        # the fp32-class plan (north_star's second parity gate, logits within 1e-4): same pipeline entry, UB_PRECISION_FP32 plan
        nf = min(64, B)
        model.b200_precision, model.b200_chunk = "fp32", nf
        f32_frames = frames_dev[:nf].contiguous()

        def step_fp32():
            model.predict_mask(f32_frames, threshold=0.5, swap_rb=True, size=(H, W), want=("mask",))

        for _ in range(3):
            step_fp32()
        n32 = max(5, args.steps // 2)
        ms32 = timed(step_fp32, n32)
        fps32 = world * nf * n32 / (ms32 / 1e3)
        tensor_flops = 3.0 * (flops_per_frame - 2.0 * H * W * 64 * 27)       # three bf16 passes per product; the stem runs on the FP32 pipes
        line["fp32_path"] = {"value": fps32, "unit": UNIT, "ms_per_step": ms32 / n32, "batch_per_gpu": nf,
                             "parity_gate": "logits within 1e-4 of the fp32 reference (tests/test_gpu_unet.py)",
                             "tensor_work_tflops_per_gpu": fps32 / world * tensor_flops / 1e12,
                             "frac_of_sustained_bf16_peak": fps32 / world * tensor_flops / 1e12 / peaks["bf16_sustained"],
                             "how": "UB_PRECISION_FP32 plan: split-bf16 (hi + lo) activations and weights, three tcgen05 K passes per product"}
        model.b200_precision, model.b200_chunk = "bf16", args.chunk
        del f32_frames
    if (H, W) == (224, 224) and not args.no_cfg5:
        # BASELINE.json configs[4]: the default-width U-Net at camera resolution (480x640 network input), batch 128 over 8 GPUs =
        # 16 frames per GPU; sources 960x1280 (the 2x down-scale of SURVEY.md 8(d) config 5: cv2 computes an exact 2x2
        # decimation as INTER_AREA) and 480x640 (resize = identity). Same kernels, different tensor maps.
        model._engines.clear()
        torch.cuda.empty_cache()
        model.b200_chunk = 16
        cfg5 = {}
        for tag, (hs5, ws5) in (("src960x1280", (960, 1280)), ("src480x640", (480, 640))):
            f5 = torch.randint(0, 256, (16, hs5, ws5, 3), dtype=torch.uint8, device=dev)

            def step5():
                model.predict_mask(f5, threshold=0.5, swap_rb=True, size=(480, 640), want=("mask",))

            for _ in range(3):
                step5()
            n5 = max(5, args.steps // 2)
            ms5 = timed(step5, n5)
            fps5 = world * 16 * n5 / (ms5 / 1e3)
            cfg5[tag] = {"value": fps5, "unit": UNIT, "ms_per_step": ms5 / n5,
                         "frac_of_sustained_bf16_peak": fps5 / world * FLOPS_PER_FRAME * (480 * 640) / (224 * 224) / 1e12 / peaks["bf16_sustained"]}
            del f5
        cfg5["config"] = (f"U-Net 480x640 bf16 inference, features {FEATURES}, 16 frames/GPU x {world} GPUs "
                          "(BASELINE.json configs[4]: batch 128 on 8 GPUs), fused resize/normalise preprocess + mask threshold")
        cfg5["flops_per_frame"] = FLOPS_PER_FRAME * (480 * 640) / (224 * 224)
        line["config5_480x640"] = cfg5
        model.b200_chunk = args.chunk
    if not args.no_train:
        del frames_dev
        model._engines.clear()
        torch.cuda.empty_cache()
        line["train"] = measure_train(dev, world, rank, args.train_batch, min(args.steps, 10), 3, timed)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and (H, W) == (224, 224):
        v, sps, threads = time_cpu_reference(32, 6, 1)      # about 10-15 s of CPU work on the box's host cores
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"32 frames/step x 6 steps ({sps * 6:.1f} s) of the oracle fp32 PyTorch CPU pipeline"}
        v1, sps1, _ = time_cpu_reference(2, 2, 1, threads=1)      # the single-thread figure as well
        line["cpu_baseline"]["one_thread"] = {"value": v1, "unit": UNIT, "cores": 1, "sample": f"2 frames/step x 2 steps ({sps1 * 2:.1f} s)"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
