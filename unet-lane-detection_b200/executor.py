"""Executor plugin with the reference's container API (src/py_utils/rknn_executor.py:4-42,
pytorch_executor.py:14-61): __init__(model_path, target=None, device_id=None), run(inputs) -> list of
numpy arrays, release(). `B200LaneInference` mirrors RKNNLaneInference (src/unet.py:14-97) on top.

The container loads a reference-format checkpoint (bare state_dict or {'model_state_dict': ...},
README.md:2205-2231) into the B200 UNet and, like the deployed RKNN graph,
applies the mean/std normalisation and the sigmoid *inside* run(): uint8 NHWC RGB frames in,
probabilities float32 [B,1,H,W] out.
"""
import numpy as np
import torch

from .unet import UNet


def _features_from_state_dict(sd):
    feats, i = [], 0
    while f"encoder_blocks.{i}.0.weight" in sd:
        feats.append(int(sd[f"encoder_blocks.{i}.0.weight"].shape[0]))
        i += 1
    if not feats:
        raise ValueError("state_dict has no encoder_blocks.*: not a reference UNet checkpoint")
    return feats, int(sd["encoder_blocks.0.0.weight"].shape[1]), int(sd["output.weight"].shape[0])


class B200_model_container:
    def __init__(self, model_path, target=None, device_id=None, output="probs"):
        if not torch.cuda.is_available():
            raise RuntimeError("B200_model_container needs a CUDA sm_100 device (no CPU fallback)")
        dev = torch.device("cuda", int(device_id) if device_id not in (None, "") else torch.cuda.current_device())
        if isinstance(model_path, UNet):
            model = model_path
        else:
            ckpt = torch.load(model_path, map_location="cpu", weights_only=True)
            sd = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
            feats, cin, cout = _features_from_state_dict(sd)
            model = UNet(cin, cout, feats)
            model.load_state_dict(sd)
        self.model = model.to(dev).eval()
        self.device = dev
        self.output = output  # "probs" (deployed graph has the sigmoid inside) or "logits"

    def run(self, inputs):
        if self.model is None:
            print("ERROR: b200 model has been released")
            return []
        if not isinstance(inputs, (list, tuple)):
            inputs = [inputs]
        frames = np.ascontiguousarray(inputs[0])
        if frames.ndim == 3:
            frames = frames[None]
        if frames.dtype != np.uint8 or frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 NHWC frames [B,H,W,3], got {frames.dtype} {frames.shape}")
        with torch.cuda.device(self.device):
            d = torch.from_numpy(frames).to(self.device, non_blocking=True)
            size = (d.shape[1], d.shape[2])
            logits, probs, _ = self.model.predict_mask(d, size=size, want=(self.output,))
            out = probs if self.output == "probs" else logits
            return [out.reshape(out.shape[0], 1, *size).cpu().numpy()]

    def release(self):
        self.model = None


class B200LaneInference:
    """RKNNLaneInference (src/unet.py:14-97) on the B200 path: predict(image, threshold) -> (mask, seconds)."""

    def __init__(self, model_path, target=None, device_id=None, input_size=(224, 224)):
        self.model = B200_model_container(model_path, target, device_id)
        self.input_size = input_size

    def predict(self, image, threshold=0.5):
        import time
        original_shape = image.shape[:2]
        t0 = time.time()
        try:
            net = self.model.model
            with torch.cuda.device(self.model.device):
                d = torch.from_numpy(np.ascontiguousarray(image)[None]).to(self.model.device)
                _, _, mask = net.predict_mask(d, threshold=threshold, size=self.input_size, want=("mask",))
                mask = mask[0].cpu().numpy()
        except Exception as e:  # reference behaviour: zero mask + elapsed time (src/unet.py:89-92)
            print(f"Inference error: {e}")
            return np.zeros(original_shape, dtype=np.uint8), time.time() - t0
        dt = time.time() - t0
        if tuple(original_shape) != tuple(self.input_size):
            import cv2  # mask up-resize back to the source size (src/unet.py:70) stays on the host for now
            mask = cv2.resize(mask, (original_shape[1], original_shape[0]))
        return mask, dt
