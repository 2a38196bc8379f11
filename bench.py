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

# one process per GPU: pin the visible device before CUDA initialises (the C-ABI library carries its own runtime)
# (the NVLink gradient exchange maps the peers' buffers, so every GPU has to stay visible: no pinning there)
PINNED = ("LOCAL_RANK" in os.environ and os.environ.get("UB_BENCH_PIN", "1") == "1"
          and ("train" not in sys.argv or "nccl" in sys.argv))
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


def physical_gpu_index():
    """Index nvidia-smi / NVML know this process's first visible GPU by (CUDA_VISIBLE_DEVICES may hold indices or UUIDs)."""
    first = os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0].strip()
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
    pre = [O.preprocess_oracle(f, (224, 224))[0] for f in frames_u8_nhwc]   # resize (identity at 224) + batch dim
    import numpy as np
    x = torch.from_numpy(O.normalize_oracle(np.concatenate(pre, 0)))
    with torch.no_grad():
        logits = model_cpu(x).numpy()
    return [O.postprocess_oracle([logits[i:i + 1]], (224, 224), threshold) for i in range(len(frames_u8_nhwc))]


def time_cpu_reference(frames_per_step, steps, warmup):
    """frames/s of the oracle pipeline on the host cores, all threads."""
    import numpy as np
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
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


def measure_train(dev, world, rank, batch, steps, warmup, timed, host_inputs=False, exchange="auto"):
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
    step = U.FusedTrainStep(net, exchange=exchange)                     # lr 1e-4, wd 1e-4, pos_weight 3 (README.md:2169-2174)
    box = {}

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
           "collective": "none" if world == 1 else (
               f"NCCL all-reduce(sum) of the flat fp32 gradient ({31037633 * 4 / 1e6:.0f} MB) per step" if step.exchange == "nccl" else
               "none: ONE kernel sums the owned gradient shard over the peers' buffers (NVLink loads), applies AdamW (ZeRO-1 "
               "sharded state) and stores the parameters to all replicas (NVLink stores); two device-side barriers per step"
               if step.exchange == "nvlink" else
               "none: ONE kernel on NVSwitch multicast addresses - multimem.ld_reduce sums the owned gradient shard inside the "
               "switch, AdamW (ZeRO-1 sharded state), multimem.st broadcasts the parameters; two device-side barriers per step"
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
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--chunk", type=int, default=256, help="frames per pass through the plan (256 = the whole batch in one pass)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: BASELINE.json headline (configs[1]); train: the training step of configs[3] as the metric")
    ap.add_argument("--train-batch", type=int, default=64, help="samples per GPU per training step (configs[3])")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "nvlink", "nvlink_pull", "nvlink_mc", "nvlink_push"],
                    help="training, N > 1: gradient exchange (auto = nvlink when the ranks can map each other's memory). nvlink: one kernel = reduce-scatter by NVLink loads + sharded AdamW + "
                         "all-gather by NVLink stores; nvlink_push: gradient atomics go to the owner GPU inside the backward kernels")
    ap.add_argument("--no-train", action="store_true", help="infer mode: skip the secondary training-step measurement")
    ap.add_argument("--hw", nargs=2, type=int, default=[224, 224], metavar=("H", "W"),
                    help="network input size; 480 640 = BASELINE.json configs[4] (camera resolution, use --batch 16 --chunk 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    B = args.batch
    H, W = args.hw
    flops_per_frame = FLOPS_PER_FRAME * (H * W) / (224 * 224)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    frames_host = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).pin_memory()
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
        sampler = ClockSampler(index=physical_gpu_index() if rank == 0 else 0)
        if rank == 0:
            sampler.start()
        tr = measure_train(dev, world, rank, args.train_batch, args.steps, args.warmup, timed, host_inputs=True,
                           exchange=args.exchange)
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
        model.predict_mask(frames_dev, threshold=0.5, size=(H, W), want=("mask",))

    def step_host():
        model.infer_host(frames_host, threshold=0.5, size=(H, W), mask_out=mask_host)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(index=physical_gpu_index() if rank == 0 else 0)
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

    # roofline of the dominant kernel family (tcgen05 implicit-GEMM convs), from per-kernel CUDA events
    peaks = load_peaks()
    x4 = torch.empty(min(args.chunk, B), H, W, 4, dtype=torch.bfloat16, device=dev).normal_()
    model.profile_layers(x4)
    rows = None
    for _ in range(3):
        r = model.profile_layers(x4)
        if rows is None:
            rows = r
        else:
            for a, b in zip(rows, r):
                a["ms"] = min(a["ms"], b["ms"])
    all_ms = sum(r["ms"] for r in rows)

    def fam(sel):
        rs = [r for r in rows if sel(r)]
        fl, ms_ = sum(r["flops"] for r in rs), sum(r["ms"] for r in rs)
        return {"launches": len(rs), "tflops": fl / (ms_ / 1e3) / 1e12 if ms_ > 0 else 0.0, "share_of_step": ms_ / all_ms,
                "flops": fl, "ms": ms_}

    # dominant kernel = conv_umma2_kernel<256>, the CTA-pair implicit GEMM (over half of the step, see profiles/r1_ncu_launches_bench_v3.csv)
    dom = fam(lambda r: r["kind"] in ("conv3x3", "convT2x2") and r["block_n"] == 256 and not r.get("halo"))
    halo64 = fam(lambda r: r.get("halo") and r["block_n"] == 64)
    halo128 = fam(lambda r: r.get("halo") and r["block_n"] == 128)
    allconv = fam(lambda r: r["kind"] in ("conv3x3", "convT2x2"))
    # DRAM bytes of the dominant kernel family from the committed ncu capture (bench.py cannot run under ncu itself):
    # sum over its launches of one pass, i.e. the same unit as flops_per_pass; only valid for the captured geometry
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_dram_traffic.json")
    if os.path.exists(tpath) and (H, W) == (224, 224):
        with open(tpath) as f:
            tj = json.load(f)
        fam_t = tj["families"].get("conv_umma2_kernel<256>")
        if fam_t and fam_t["launches"] == dom["launches"] and tj.get("chunk") == int(x4.shape[0]):
            traffic = fam_t["dram_read_bytes"] + fam_t["dram_write_bytes"]
            traffic_src = "profiles/r1_dram_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, summed over the family's launches of one pass)"
    roofline = {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": dom["tflops"] / peaks["bf16_sustained"], "traffic": traffic, "traffic_unit": "bytes per pass", "traffic_source": traffic_src,
                "kernel": f"conv_umma2_kernel<256> (cta_group::2; {dom['launches']} launches per pass: 3x3 convs with Cout>=256 + 4 ConvT)",
                "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "share_of_step": dom["share_of_step"], "flops_per_pass": dom["flops"], "ms_per_pass": dom["ms"],
                "timing": f"CUDA events around every kernel of one {int(x4.shape[0])}-frame pass on the launching stream, min of 3",
                "traffic_note": "reads equal the layers' input bytes exactly (no re-reads; e.g. enc0.conv1 reads 822.5 MB = 128x224x224x64 bf16); "
                                "ncu --set full captures in profiles/r1_ncu_full_*.txt",
                "other_kernels": {"conv_halo2_kernel<64>": halo64, "conv_halo2_kernel<128>": halo128, "all_tensor_core_convs": allconv},
                "whole_net_frac_of_peak": (value / world) * flops_per_frame / 1e12 / peaks["bf16_sustained"]}
    if args.layers_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
        with open(args.layers_out, "w") as f:
            json.dump({"chunk": int(x4.shape[0]), "rows": rows, "peaks": peaks}, f, indent=1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"U-Net {H}x{W} bf16 inference, features {FEATURES}, batch {B}/GPU, fused preprocess + mask threshold "
                               + ("(BASELINE.json configs[1])" if (H, W) == (224, 224) else "(BASELINE.json configs[4] geometry)"),
                   "batch_per_gpu": B, "chunk": args.chunk,
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2_policy": f"inputs+activations >> L2: {B * H * W * 3 / 1e6:.0f} MB frames and "
                                f"{min(args.chunk, B) * 64.1 * H * W / (224 * 224):.0f} MB activations per chunk vs 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * H * W * 3, "d2h_bytes_per_step": B * H * W,
                "ms_per_step": ms_host / args.steps, "api": "UNet.infer_host -> unet_b200_infer_u8_host_stream (pinned host buffers, copies overlapped with compute)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if (H, W) != (224, 224):
        line["metric"] = f"unet{H}x{W}_inference_frames_per_sec"
        args.no_train = True
    if not args.no_train:
        del frames_dev
        model._engines.clear()
        torch.cuda.empty_cache()
        line["train"] = measure_train(dev, world, rank, args.train_batch, min(args.steps, 10), 3, timed)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and (H, W) == (224, 224):
        v, sps, threads = time_cpu_reference(32, 6, 1)      # about 10-15 s of CPU work on the box's host cores
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"32 frames/step x 6 steps ({sps * 6:.1f} s) of the oracle fp32 PyTorch CPU pipeline"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
