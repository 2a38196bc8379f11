"""GPU check of the training step against the CPU oracle (fp32 PyTorch): logits, loss, every parameter gradient,
BN running statistics, and one AdamW step. Prints per-tensor relative errors.
    python tools/train_check.py [features...] --hw 32 32 --batch 4
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_lane_detection_b200 as U  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("features", nargs="*", type=int, default=[64, 128])
    ap.add_argument("--hw", nargs=2, type=int, default=[32, 32])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--fused", action="store_true")
    args = ap.parse_args()
    feats, (H, W), B = args.features, args.hw, args.batch
    torch.manual_seed(0)
    ref = O.UNetOracle(3, 1, feats).train()
    O.randomize_bn_(ref, seed=1)
    O.scale_head_(ref, 10.0)
    net = U.UNet(3, 1, feats)
    net.load_state_dict(ref.state_dict())
    net = net.cuda().train()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, 3, H, W, generator=g)
    y = (torch.rand(B, 1, H, W, generator=g) < 0.085).float()
    crit = O.BCEDiceLossOracle(0.5, 0.5, pos_weight=torch.tensor([3.0]), smooth=1e-6)

    # oracle (fp32) and the same network with the B200 rounding points (bf16-emulated)
    import copy
    emu = copy.deepcopy(ref)
    out_ref = ref(x)
    loss_ref, bce_ref, dice_ref = crit(out_ref, y)
    loss_ref.backward()
    out_emu = O.forward_train_bf16_emulated(emu, x)
    crit(out_emu, y)[0].backward()

    if args.fused:
        step = U.FusedTrainStep(net)
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-4, weight_decay=1e-4)
        losses = step.step(x.cuda(), y.cuda())
        torch.cuda.synchronize()
        print("loss  b200", losses.tolist(), " oracle", [loss_ref.item(), bce_ref.item(), dice_ref.item()])
        grads = step.last_grads.cpu()
        off = 0
        worst = 0.0
        for (n, p) in ref.named_parameters():
            k = p.numel()
            r = rel(grads[off:off + k].view(p.shape), p.grad)
            worst = max(worst, r)
            off += k
        print(f"worst grad rel err {worst:.4f}")
        before = {n: p.detach().clone() for n, p in ref.named_parameters()}
        opt.step()
        worst_upd = 0.0
        for (n, p), q in zip(ref.named_parameters(), net.parameters()):
            du_ref = p.detach() - before[n]
            du = q.detach().cpu() - before[n]
            worst_upd = max(worst_upd, rel(du, du_ref))
        print(f"worst AdamW update rel err {worst_upd:.4f}")
        return

    out = net(x.cuda())
    print("logits: max|d| %.4e  rel %.4e  (range %.3f)" % ((out.detach().cpu() - out_ref.detach()).abs().max().item(),
                                                          rel(out.detach().cpu(), out_ref.detach()), out_ref.abs().max().item()))
    crit_gpu = O.BCEDiceLossOracle(0.5, 0.5, pos_weight=torch.tensor([3.0]).cuda(), smooth=1e-6)
    loss, bce, dice = crit_gpu(out, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print("loss  b200 %.6f %.6f %.6f   oracle %.6f %.6f %.6f" % (loss.item(), bce.item(), dice.item(), loss_ref.item(),
                                                                bce_ref.item(), dice_ref.item()))
    # fused loss kernel on the oracle's logits
    l3, dz = U.bce_dice_loss(out_ref.detach().cuda(), y.cuda())
    zz = out_ref.detach().clone().requires_grad_(True)
    crit(zz, y)[0].backward()
    print("fused loss kernel:", l3.tolist(), " dz rel %.3e" % rel(dz.cpu().reshape(-1), zz.grad.reshape(-1)))
    print("logits vs emulated: max|d| %.4e" % (out.detach().cpu() - out_emu.detach()).abs().max().item())
    for (n, p), (n2, q), (_, e) in zip(ref.named_parameters(), net.named_parameters(), emu.named_parameters()):
        assert n == n2
        gq = q.grad.detach().cpu()
        cos = torch.nn.functional.cosine_similarity(gq.reshape(-1), p.grad.reshape(-1), dim=0).item()
        print(f"{n:34s} {str(tuple(p.shape)):22s} vs fp32: rel {rel(gq, p.grad):.4f} cos {cos:.5f} | emu-vs-fp32 rel "
              f"{rel(e.grad, p.grad):.4f} | vs emu: rel {rel(gq, e.grad):.4f} |g| {p.grad.norm().item():.3e}")
    for (n, b), (n2, c) in zip(ref.named_buffers(), net.named_buffers()):
        if "num_batches" in n:
            assert int(b) == int(c), (n, int(b), int(c))
            continue
        r = rel(c.detach().cpu(), b)
        if r > 1e-3:
            print(f"buffer {n}: rel {r:.4e}")
    print("buffers checked")


if __name__ == "__main__":
    main()
