"""Per-frame latency of the executor shim (what src/unet.py's RKNNLaneInference calls once per camera frame):
B200_model_container.run([uint8 NHWC (1,224,224,3)]) -> [float (1,1,224,224)], host numpy in, host numpy out."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402

import tempfile  # noqa: E402

torch.manual_seed(0)
# the reference's contract: the container is built from a checkpoint PATH (weights fixed from then on); a container around a
# live nn.Module instead re-checks the 118 tensors' versions on every call (+ ~0.1 ms)
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "unet.pth")
    torch.save({"model_state_dict": U.UNet(3, 1, [64, 128, 256, 512]).state_dict()}, path)
    box = U.B200_model_container(path)
frame = np.random.default_rng(0).integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
for _ in range(20):
    box.run([frame])
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    out = box.run([frame])
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"run() latency, batch 1 @224x224 (host in -> host out): median {ts[100]:.3f} ms, p10 {ts[20]:.3f}, p90 {ts[180]:.3f}; "
      f"output {out[0].shape} {out[0].dtype}")
# where the time goes: the captured pass alone (GPU time between two events, replays back to back) and replay + synchronize
graph = next(iter(box._graphs.values()))[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    graph.replay()
e1.record()
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 200
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    graph.replay()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"captured pass alone: {gpu_ms:.3f} ms of GPU time per replay (back to back); replay + synchronize from the host: median {ts[100]:.3f} ms")

# the ROS callback as INTEGRATION.md wires it: B200LanePipeline.process(bgr 480x640 frame) -> uint8 mask 685x1055 (IPM warp,
# resize, network, threshold and mask up-resize on the GPU; one captured pass per call)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ipm.npz"))
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "unet.pth")
    torch.save({"model_state_dict": U.UNet(3, 1, [64, 128, 256, 512]).state_dict()}, path)
    pipe = U.B200LanePipeline(path, g["M"], threshold=0.5)
cam = np.random.default_rng(1).integers(0, 256, (480, 640, 3), dtype=np.uint8)
for _ in range(20):
    pipe.process(cam)
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    m = pipe.process(cam)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"B200LanePipeline.process, one 480x640 camera frame -> mask {m.shape}: median {ts[100]:.3f} ms, p10 {ts[20]:.3f}, p90 {ts[180]:.3f}")

# RKNNLaneInference.predict's mirror: image -> mask at the image's own size (resize to 224 x 224 inside, mask resized back)
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "unet.pth")
    torch.save({"model_state_dict": U.UNet(3, 1, [64, 128, 256, 512]).state_dict()}, path)
    inf = U.B200LaneInference(path)
for _ in range(20):
    inf.predict(cam, 0.5)
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    m, _ = inf.predict(cam, 0.5)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"B200LaneInference.predict, one 480x640 image -> mask {m.shape}: median {ts[100]:.3f} ms, p10 {ts[20]:.3f}, p90 {ts[180]:.3f}")
