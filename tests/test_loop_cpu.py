"""CPU: host logic of the training loop (loop.py) against PyTorch's own scheduler / optimizer structures."""
import torch

import unet_lane_detection_b200 as U
from unet_lane_detection_b200.loop import EarlyStopping


def test_cosine_warm_restarts_matches_torch_scheduler():
    for t0, tm in ((10, 2), (10, 1), (3, 3), (1, 2)):           # README.md:2177 uses (10, 2)
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=1e-4)
        sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=t0, T_mult=tm)
        for epoch in range(150):
            want = opt.param_groups[0]["lr"]
            got = U.cosine_warm_restarts_lr(1e-4, epoch, T_0=t0, T_mult=tm)
            assert abs(got - want) <= 1e-18 + 1e-12 * want, (t0, tm, epoch, got, want)
            opt.step()
            sch.step()


def test_early_stopping_follows_reference_rule():
    """README.md:2205-2221: strict improvement resets the counter; stop when counter >= patience."""
    es = EarlyStopping(patience=3)
    seq = [0.5, 0.6, 0.6, 0.59, 0.61, 0.1, 0.2, 0.3]
    out = [es.update(v) for v in seq]
    assert out == [(True, False), (True, False), (False, False), (False, False), (True, False), (False, False), (False, False),
                   (False, True)]
    assert es.best == 0.61


def test_optimizer_state_dict_has_torch_adamw_structure():
    m = U.UNet(3, 1, [64, 128])
    step = U.FusedTrainStep(m, lr=3e-4, weight_decay=1e-2)
    mine = step.state_dict()
    ref = torch.optim.AdamW(m.parameters(), lr=3e-4, weight_decay=1e-2).state_dict()
    assert mine["state"] == {} and ref["state"] == {}
    g, r = mine["param_groups"][0], ref["param_groups"][0]
    assert g["params"] == r["params"]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert g[k] == r[k], k
    torch.optim.AdamW(m.parameters()).load_state_dict(mine)       # loads without complaint
