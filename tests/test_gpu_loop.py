"""GPU (B200): the README-style training loop (loop.py: train_one_epoch / validate / fit) on a synthetic lane dataset,
fused validation metrics against the oracle, checkpoint files and optimizer-state interchange."""
import os

import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import unet_lane_detection_b200 as mod
    return mod


def lane_batches(n_batches, batch, hw, seed):
    """Synthetic 'lanes': two bright slanted stripes on a noisy road; mask = stripe pixels."""
    g = torch.Generator().manual_seed(seed)
    H, W = hw
    ys = torch.arange(H).view(H, 1).float()
    xs = torch.arange(W).view(1, W).float()
    out = []
    for _ in range(n_batches):
        imgs, masks = [], []
        for _ in range(batch):
            x0 = torch.rand(2, generator=g) * W * 0.6 + W * 0.2
            slope = (torch.rand(2, generator=g) - 0.5) * 0.8
            m = torch.zeros(H, W)
            for k in range(2):
                m = torch.maximum(m, ((xs - (x0[k] + slope[k] * ys)).abs() < 2.0).float())
            img = torch.randn(3, H, W, generator=g) * 0.5 + m * 2.0
            imgs.append(img)
            masks.append(m[None])
        out.append((torch.stack(imgs), torch.stack(masks)))
    return out


def test_validation_metrics_match_oracle(U):
    g = torch.Generator().manual_seed(1)
    z = torch.randn(6, 1, 64, 64, generator=g) * 3
    t = (torch.rand(6, 1, 64, 64, generator=g) < 0.1).float()
    crit = O.BCEDiceLossOracle(0.5, 0.5, pos_weight=torch.tensor([3.0]), smooth=1e-6)
    tot, bce, dice = crit(z, t)
    score = O.compute_dice_oracle(torch.sigmoid(z) > 0.5, t)              # README.md:2103-2104, 2115-2120
    got = U.validation_metrics(z.cuda(), t.cuda()).cpu()
    want = torch.tensor([tot.item(), bce.item(), dice.item(), score])
    assert (got - want).abs().max().item() <= 2e-6


def test_fit_trains_checkpoints_and_resumes(U, tmp_path):
    torch.manual_seed(0)
    model = U.UNet(3, 1, [64, 128]).cuda()
    train = lane_batches(6, 8, (64, 64), seed=1)
    val = lane_batches(2, 8, (64, 64), seed=2)
    cfg = {"epochs": 10, "learning_rate": 2e-3, "weight_decay": 1e-4, "patience": 15, "seed": 42, "save_dir": str(tmp_path)}
    U.fit(model, train, val, cfg)
    hist = model.b200_history
    assert len(hist) == 10
    assert abs(hist[3]["lr"] - U.cosine_warm_restarts_lr(2e-3, 3)) < 1e-12          # scheduler stepped once per epoch
    assert hist[-1]["train_loss"] < 0.6 * hist[0]["train_loss"]
    assert max(h["val_dice"] for h in hist) > 0.6                                   # it learns the stripes
    # files of README.md:2205-2231
    best = torch.load(os.path.join(tmp_path, "best_model.pth"), weights_only=False)
    assert set(best) == {"epoch", "model_state_dict", "optimizer_state_dict", "best_dice"}
    assert os.path.exists(os.path.join(tmp_path, "checkpoint_epoch10.pth")) and os.path.exists(os.path.join(tmp_path, "last_model.pth"))
    # the checkpoint loads into the reference architecture and reproduces the validation Dice in fp32 on the CPU
    ref = O.UNetOracle(3, 1, [64, 128])
    ref.load_state_dict(torch.load(os.path.join(tmp_path, "last_model.pth")))
    ref.eval()
    with torch.no_grad():
        d_ref = sum(O.compute_dice_oracle(torch.sigmoid(ref(x)) > 0.5, y) for x, y in val) / len(val)
    assert abs(d_ref - hist[-1]["val_dice"]) < 0.02
    # optimizer state interchanges with torch.optim.AdamW (README.md:2173) in both directions
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    opt.load_state_dict(best["optimizer_state_dict"])
    step = U.FusedTrainStep(model)
    step.load_state_dict(opt.state_dict())
    assert step.step_count == int(best["optimizer_state_dict"]["state"][0]["step"])
    model.train()
    losses = step.step(train[0][0].cuda(), train[0][1].cuda())
    assert torch.isfinite(losses).all()
