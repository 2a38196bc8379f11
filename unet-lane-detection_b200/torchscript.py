"""TorchScript compatibility: `torch.jit.trace(model, example)` of the B200 UNet serialises, and the
reference's `Torch_model_container` (src/py_utils/pytorch_executor.py:14-61: torch.jit.load -> model(torch.tensor(np)) ->
.cpu().numpy()) runs the file, provided this package has been imported in the loading process (the import registers the
operator; INTEGRATION.md shows the one line).

The forward of the B200 module is a C-ABI call, which a tracer cannot see. Under tracing the module therefore emits ONE
dispatcher operator, `unet_b200::infer`, whose arguments are the input, the model's whole state as one flat fp32 constant and
the constructor arguments; the operator's implementation (below, Python, registered through torch.library) rebuilds a frozen
`UNet` around that state once per process and runs the same kernels as the eager module. No PyTorch compute, no CPU fallback:
the operator raises without a CUDA sm_100 device.

mode 0  x float NCHW [B,in_channels,H,W]            -> logits [B,out_channels,H,W] (what UNet.forward returns)
mode 1  x uint8 NHWC RGB [B,H,W,3] (the lane node's frames, src/unet.py:24-42)
                                                    -> probabilities float32 [B,1,H,W]: normalisation and sigmoid inside the
                                                       graph, like the deployed RKNN blob
mode 2  x uint8 NHWC RGB                            -> logits float32 [B,1,H,W]
The result lives on the device of `x` (Torch_model_container feeds CPU tensors and calls .cpu() on the result).
"""
from collections import OrderedDict
from typing import List

import torch

_LIB = torch.library.Library("unet_b200", "DEF")
_LIB.define("infer(Tensor x, Tensor state, int[] features, int in_channels, int out_channels, str precision, int mode) -> Tensor")

_NETS = OrderedDict()   # (state ptr, version, device, features, ...) -> frozen UNet; a handful of models per process at most
_MAX_NETS = 4


def state_keys(model):
    """state_dict entries that make up the flat state, in order (the integer BatchNorm step counters are left out)."""
    return [k for k in model.state_dict().keys() if not k.endswith("num_batches_tracked")]


def flat_state(model):
    """One contiguous fp32 tensor holding every parameter and BatchNorm buffer of `model` in state_dict order."""
    sd = model.state_dict()
    return torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in state_keys(model)]).contiguous()


def _net_for(state, features, in_channels, out_channels, precision):
    from .unet import UNet
    if not torch.cuda.is_available():
        raise RuntimeError("unet_b200::infer needs a CUDA sm_100 device (no CPU fallback)")
    dev = state.device if state.is_cuda else torch.device("cuda", torch.cuda.current_device())
    key = (state.data_ptr(), state._version, str(dev), tuple(features), in_channels, out_channels, precision)
    net = _NETS.get(key)
    if net is not None:
        _NETS.move_to_end(key)
        return net
    net = UNet(in_channels, out_channels, list(features))
    sd, off = net.state_dict(), 0
    src = state.detach().to("cpu", torch.float32)
    for k in state_keys(net):
        n = sd[k].numel()
        if off + n > src.numel():
            raise RuntimeError("unet_b200::infer: the state tensor is shorter than the model it describes")
        sd[k].copy_(src[off:off + n].view(sd[k].shape))
        off += n
    if off != src.numel():
        raise RuntimeError(f"unet_b200::infer: state holds {src.numel()} values, the model {off}")
    net = net.to(dev).eval()
    net.b200_precision = precision
    net.b200_frozen = True
    while len(_NETS) >= _MAX_NETS:
        _NETS.popitem(last=False)
    _NETS[key] = net
    return net


def _infer(x, state, features: List[int], in_channels: int, out_channels: int, precision: str, mode: int):
    net = _net_for(state, features, in_channels, out_channels, precision)
    dev = next(net.parameters()).device
    with torch.no_grad(), torch.cuda.device(dev):
        if mode == 0:
            y = net(x.to(dev))
        elif mode in (1, 2):
            if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
                raise RuntimeError(f"unet_b200::infer mode {mode} takes uint8 NHWC frames [B,H,W,3], got {x.dtype} {tuple(x.shape)}")
            frames = x.to(dev).contiguous()
            want = "probs" if mode == 1 else "logits"
            logits, probs, _ = net.predict_mask(frames, size=(int(x.shape[1]), int(x.shape[2])), want=(want,))
            out = probs if mode == 1 else logits
            y = out.reshape(x.shape[0], -1, x.shape[1], x.shape[2])
        else:
            raise RuntimeError(f"unet_b200::infer: unknown mode {mode}")
    return y if x.is_cuda else y.to(x.device)


_LIB.impl("infer", _infer, "CompositeExplicitAutograd")


def traced_forward(model, x, mode=0):
    """What UNet.forward does while torch.jit.trace is recording: the state becomes a constant of the graph (built with
    recording paused, so the 118 tensors do not turn into 118 graph inputs plus a concatenation per call)."""
    ts = torch._C._get_tracing_state()
    torch._C._set_tracing_state(None)
    try:
        state = flat_state(model).to(next(model.parameters()).device)
    finally:
        torch._C._set_tracing_state(ts)
    return torch.ops.unet_b200.infer(x, state, list(model.features), int(model.in_channels), int(model.out_channels),
                                     str(model.b200_precision), int(mode))


class LaneGraph(torch.nn.Module):
    """The deployed graph's contract around the UNet: uint8 NHWC RGB frames in, probabilities
    [B,1,H,W] out (logits with sigmoid=False). Exists to be traced: `export_torchscript`."""

    def __init__(self, model, sigmoid=True):
        super().__init__()
        self.model = model
        self.mode = 1 if sigmoid else 2

    def forward(self, frames_u8):
        if torch.jit.is_tracing():
            return traced_forward(self.model, frames_u8, self.mode)
        return torch.ops.unet_b200.infer(frames_u8, flat_state(self.model).to(next(self.model.parameters()).device),
                                         list(self.model.features), int(self.model.in_channels), int(self.model.out_channels),
                                         str(self.model.b200_precision), self.mode)


def export_torchscript(model, path, input_kind="float_nchw", sigmoid=True, example=None):
    """Write a TorchScript file `Torch_model_container(path)` can run (with this package imported).
    input_kind "float_nchw": the module itself (float NCHW -> logits); "uint8_nhwc": LaneGraph (frames -> probabilities)."""
    model = model.eval()
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("export_torchscript: move the model to a CUDA device first (tracing runs the B200 forward once)")
    if input_kind == "float_nchw":
        ex = example if example is not None else torch.zeros(1, model.in_channels, 224, 224, device=dev)
        traced = torch.jit.trace(model, ex, check_trace=False)
    elif input_kind == "uint8_nhwc":
        ex = example if example is not None else torch.zeros(1, 224, 224, 3, dtype=torch.uint8, device=dev)
        traced = torch.jit.trace(LaneGraph(model, sigmoid=sigmoid), ex, check_trace=False)
    else:
        raise ValueError("input_kind must be 'float_nchw' or 'uint8_nhwc'")
    torch.jit.save(traced, path)
    return traced
