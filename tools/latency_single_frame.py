"""Per-frame latency of the executor shim (what src/unet.py's RKNNLaneInference calls once per camera frame):
B200_model_container.run([uint8 NHWC (1,224,224,3)]) -> [float (1,1,224,224)], host numpy in, host numpy out."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_lane_detection_b200 as U  # noqa: E402

torch.manual_seed(0)
box = U.B200_model_container(U.UNet(3, 1, [64, 128, 256, 512]))
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
