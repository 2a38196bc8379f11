"""Host-side training loop around the B200 step, mirroring the reference's listing function by function
(README.md:2060-2084 train_one_epoch, 2086-2112 validate, 2115-2120 compute_dice, 2125-2234 train): same epoch
structure, same learning-rate schedule (CosineAnnealingWarmRestarts(T_0=10, T_mult=2), stepped once per epoch,
README.md:2177, 2198), same checkpoint files and dict keys (README.md:2205-2231), same early stopping.

The data pipeline (LaneDataset + albumentations, README.md:1996-2055) is out of scope: any iterable of
(images float NCHW, masks) batches works, on the host (pinned memory recommended) or already on the device.
All arithmetic on tensors - forward, loss, backward, optimizer, validation metrics - runs in libunet_b200.so.
"""
import logging
import math
import os

import torch

from ._lib import check, lib
from .training import FusedTrainStep


def cosine_warm_restarts_lr(base_lr, epoch, T_0=10, T_mult=2, eta_min=0.0):
    """Learning rate of torch.optim.lr_scheduler.CosineAnnealingWarmRestarts after `epoch` calls of scheduler.step()
    (epoch 0 = the initial rate). README.md:2177 uses T_0=10, T_mult=2."""
    if T_mult == 1:
        t_cur, t_i = epoch % T_0, T_0
    else:
        n = int(math.log(epoch / T_0 * (T_mult - 1) + 1, T_mult)) if epoch >= T_0 else 0
        t_cur = epoch - T_0 * (T_mult ** n - 1) // (T_mult - 1)
        t_i = T_0 * T_mult ** n
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


def validation_metrics(logits, target, pos_weight=3.0, bce_weight=0.5, dice_weight=0.5, smooth=1e-6, threshold=0.5):
    """One fused pass (unet_b200_validation_metrics): device tensor [total loss, bce, dice loss, Dice score of the
    thresholded prediction] = criterion(outputs, masks) and compute_dice(sigmoid(outputs) > 0.5, masks)
    (README.md:2099-2104, 2115-2120)."""
    z = logits.contiguous().to(torch.float32)
    t = target.contiguous().to(torch.float32)
    if not z.is_cuda:
        raise RuntimeError("validation_metrics (B200): CUDA tensors only - there is no CPU fallback")
    if z.numel() != t.numel():
        raise ValueError("logits and target must have the same number of elements")
    scratch = torch.empty(6, dtype=torch.float64, device=z.device)
    out = torch.empty(4, dtype=torch.float32, device=z.device)
    check(lib.unet_b200_validation_metrics(z.data_ptr(), t.data_ptr(), z.numel(), float(pos_weight), float(bce_weight),
                                           float(dice_weight), float(smooth), float(threshold), scratch.data_ptr(),
                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


def train_one_epoch(step: FusedTrainStep, dataloader, device="cuda"):
    """README.md:2060-2084: one pass over the loader, returns running_loss / len(dataloader). The per-batch losses are
    accumulated on the device; the host reads ONE number per epoch instead of loss.item() per batch."""
    step.model.train()
    total = torch.zeros((), dtype=torch.float32, device=device)
    n = 0
    for images, masks in dataloader:
        images = images.to(device, non_blocking=True)
        masks = masks.to(device, non_blocking=True)
        losses = step.step(images, masks)
        total += losses[0]
        n += 1
    return float(total.item()) / max(n, 1)


@torch.no_grad()
def validate(model, dataloader, device="cuda", pos_weight=3.0, bce_weight=0.5, dice_weight=0.5, smooth=1e-6, process_group=None):
    """README.md:2086-2112: eval-mode forward (folded BatchNorm), criterion and thresholded Dice per batch, batch means.
    Data-parallel (torch.distributed initialised): every rank validates ITS loader and the sums are all-reduced, so all
    ranks return the same two numbers (the mean over all batches of all ranks) and take the same best-model / early-stop
    decisions."""
    model.eval()
    acc = torch.zeros(5, dtype=torch.float32, device=device)   # 4 metric sums + the batch count
    for images, masks in dataloader:
        images = images.to(device, non_blocking=True)
        masks = masks.to(device, non_blocking=True)
        outputs = model(images)
        acc[:4] += validation_metrics(outputs, masks, pos_weight, bce_weight, dice_weight, smooth)
        acc[4] += 1
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(process_group) > 1:
        torch.distributed.all_reduce(acc, op=torch.distributed.ReduceOp.SUM, group=process_group)
    acc = acc.tolist()
    n = max(acc[4], 1.0)
    return acc[0] / n, acc[3] / n


class EarlyStopping:
    """best-by-validation-Dice bookkeeping of README.md:2205-2221."""

    def __init__(self, patience):
        self.patience, self.best, self.counter = patience, 0.0, 0

    def update(self, val_dice):
        """Returns (is_best, should_stop)."""
        if val_dice > self.best:
            self.best, self.counter = val_dice, 0
            return True, False
        self.counter += 1
        return False, self.counter >= self.patience


def fit(model, train_loader, val_loader, config, process_group=None, logger=None):
    """README.md:2125-2234 `train(config)` from the model onwards. config keys as in README.md:2240-2250:
    epochs, learning_rate, weight_decay, patience, save_dir (+ optional seed). Returns the model.
    Files written: best_model.pth {'epoch','model_state_dict','optimizer_state_dict','best_dice'},
    checkpoint_epoch{N}.pth {'epoch','model_state_dict'} every 10 epochs, last_model.pth (bare state_dict)."""
    logger = logger or logging.getLogger(__name__)
    device = next(model.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("fit (B200): move the model to a CUDA device first - there is no CPU fallback")
    if "seed" in config:
        torch.manual_seed(config["seed"])
        torch.cuda.manual_seed_all(config["seed"])
    base_lr = config["learning_rate"]
    step = FusedTrainStep(model, lr=base_lr, weight_decay=config["weight_decay"], process_group=process_group)
    stopper = EarlyStopping(config["patience"])
    distributed = torch.distributed.is_available() and torch.distributed.is_initialized() and \
        torch.distributed.get_world_size(process_group) > 1
    rank0 = (not distributed) or torch.distributed.get_rank(process_group) == 0
    os.makedirs(config["save_dir"], exist_ok=True)
    history = []
    for epoch in range(1, config["epochs"] + 1):
        step.lr = cosine_warm_restarts_lr(base_lr, epoch - 1)          # scheduler.step() ran epoch-1 times so far
        logger.info("Epoch %d/%d  Learning Rate: %.6f", epoch, config["epochs"], step.lr)
        train_loss = train_one_epoch(step, train_loader, device)
        if distributed:
            # BatchNorm running statistics are per replica (batch statistics of the replica's own samples, as a
            # single-device reference run would see them): evaluate and checkpoint rank 0's, on every rank
            src = torch.distributed.get_global_rank(process_group, 0) if process_group is not None else 0
            for buf in model.buffers():
                torch.distributed.broadcast(buf, src=src, group=process_group)
            model._b200_epoch += 1
        val_loss, val_dice = validate(model, val_loader, device, process_group=process_group,
                                      **{k: step.loss_cfg[k] for k in step.loss_cfg})
        history.append({"epoch": epoch, "lr": step.lr, "train_loss": train_loss, "val_loss": val_loss, "val_dice": val_dice})
        logger.info("Train Loss: %.4f  Val Loss: %.4f, Val Dice: %.4f", train_loss, val_loss, val_dice)
        # val_dice is identical on every rank (validate all-reduces), so is_best / stop are too: all ranks leave the loop together
        is_best, stop = stopper.update(val_dice)
        if is_best:
            opt_sd = step.state_dict()       # collective when the optimizer state is sharded over the replicas: every rank calls it
            if rank0:
                torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt_sd,
                            "best_dice": stopper.best}, os.path.join(config["save_dir"], "best_model.pth"))
        if stop:
            logger.info("Early stopping at epoch %d", epoch)
            break
        if epoch % 10 == 0 and rank0:
            torch.save({"epoch": epoch, "model_state_dict": model.state_dict()},
                       os.path.join(config["save_dir"], f"checkpoint_epoch{epoch}.pth"))
    if rank0:
        torch.save(model.state_dict(), os.path.join(config["save_dir"], "last_model.pth"))
    logger.info("Training completed! Best Dice: %.4f", stopper.best)
    model.b200_history = history
    return model
