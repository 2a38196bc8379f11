"""GPU x2 (skipped on a single-GPU box): data-parallel FusedTrainStep. With the SAME batch on both ranks the
summed-and-rescaled gradient equals the single-GPU gradient, so both replicas must end up with (numerically) the
parameters a single-GPU step produces, and stay identical to each other. Also: fit() on two ranks (rank-symmetric
collectives, same stop decision), optimizer-state save / resume with the sharded (NVLink) state, and two executors on two
devices inside one process (per-device library state)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXCHANGES = {   # name -> (exchange, overlap, bucket_elems)
    "nccl": ("nccl", True, 20000), "nccl_one_bucket": ("nccl", False, 1 << 21), "nvlink": ("nvlink_pull", True, 20000),
    "nvlink_one_bucket": ("nvlink_pull", False, 1 << 21), "nvlink_mc": ("nvlink_mc", True, 20000), "nvlink_push": ("nvlink_push", True, 1 << 21),
}


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _finish(q, payload):
    q.put(payload)
    q.close()
    q.join_thread()
    try:
        dist.destroy_process_group()
    finally:
        os._exit(0)


def _exchange_case(U, name, rank, world):
    exchange, overlap, bucket = EXCHANGES[name]
    torch.manual_seed(rank)          # replicas start DIFFERENT: the step must broadcast rank 0's parameters and buffers
    net = U.UNet(3, 1, [64, 128]).cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g).cuda()
    y = (torch.rand(8, 1, 32, 32, generator=g) < 0.2).float().cuda()
    step = U.FusedTrainStep(net, lr=1e-3, exchange=exchange, overlap=overlap, bucket_elems=bucket)
    gsum, nb = None, 0
    for i in range(4):                # step 0 eager, step 1 captures the graph, steps 2-3 replay it
        try:
            losses = step.step(x, y)
        except RuntimeError as e:
            if "NVLS multicast" not in str(e):
                raise
            return ("skip", str(e))
        if i == 0:   # summed gradient of the first step (parameters still identical to the single-GPU run)
            nb = len(step.buckets or [])
            if step.nvlink is None:
                gsum = step.last_grads.clone().cpu()        # all-reduced in place
            elif step.nvlink.mode != "push":
                full = torch.zeros(step.nvlink.n, device="cuda")
                for a, b, t in step.nvlink.reduced_parts():  # the sum exactly as the exchange kernel forms it
                    full[a:b] = t
                dist.all_reduce(full)                        # parts are disjoint: gather
                gsum = full.cpu()
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    return (float((others[0] - others[1]).abs().max()), flat.cpu(), losses.cpu(), gsum, nb, step.exchange)


def _worker(rank, world, port, q):
    """Every exchange variant in ONE pair of processes (process start-up and NCCL initialisation dominate otherwise)."""
    _init(rank, world, port)
    import unet_lane_detection_b200 as U
    out = {}
    for name in EXCHANGES:
        try:
            out[name] = _exchange_case(U, name, rank, world)
        except Exception as e:  # noqa: BLE001  (reported per variant; a CUDA error poisons the rest, which then fail too)
            out[name] = ("error", f"{type(e).__name__}: {e}"[:2000])
    _finish(q, (rank, out))


def _spawn(target, args, n=2, timeout=150):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, n) + args[:1] + (q,) + args[1:]) for r in range(n)]
    for p in procs:
        p.start()
    try:
        res = sorted([q.get(timeout=timeout) for _ in procs], key=lambda r: r[0])
    finally:                      # a worker that died leaves its peer waiting in a collective: never wait for it
        for p in procs:
            p.join(timeout=10 if p.exitcode is None else 1)
            if p.is_alive():
                p.kill()
    return res


def _port(salt):
    return 33500 + (os.getpid() % 2000) + salt


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_two_gpu_step_matches_single_gpu():
    """nccl: per gradient bucket an all-reduce + AdamW on a side stream, overlapped with the rest of the backward;
    nvlink: per bucket ONE kernel sums its part over the peers' buffers (NVLink loads), applies the sharded AdamW and stores
    the new parameters to all replicas; nvlink_mc: the same on NVSwitch multicast addresses; *_one_bucket: the un-overlapped
    form (one exchange after the whole backward); nvlink_push: gradient atomics routed to the owner inside the backward."""
    _need_two()
    res = _spawn(_worker, (_port(0),), timeout=400)
    # single-GPU reference in this process (rank 0's initialisation: seed 0)
    import unet_lane_detection_b200 as U
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128]).cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g).cuda()
    y = (torch.rand(8, 1, 32, 32, generator=g) < 0.2).float().cuda()
    step = U.FusedTrainStep(net, lr=1e-3)
    g1 = None
    for i in range(4):
        losses = step.step(x, y)
        if i == 0:
            g1 = step.last_grads.clone().cpu()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    ran = []
    for name in EXCHANGES:
        r0 = res[0][1][name]
        if r0[0] == "skip":
            continue
        assert r0[0] != "error", (name, r0[1], res[1][1][name])
        spread, flat_n, losses_n, gsum, nb, used = r0
        ran.append((name, used, nb))
        if EXCHANGES[name][1] and name != "nvlink_push":
            assert nb >= 3, f"{name}: expected several gradient buckets, got {nb}"
        # replicas stay in lock step up to the run-to-run noise of fp32 atomics (identical all-reduced gradients are applied
        # to identical parameters; the only divergence is each replica's own forward nondeterminism in later steps)
        assert spread < 2 * 4 * 1e-3 + 2e-4, (name, spread)
        # the exchanged gradient of step 1 is world x the single-GPU gradient (same batch on both replicas): this is the
        # parity check of the reduce-scatter inside the exchange kernel (nvlink) / of the bucketed all-reduce (nccl)
        if gsum is not None:
            gerr = (gsum - 2.0 * g1).abs().max().item() / (2.0 * g1.abs().max().item())
            assert gerr < 2e-3, (name, gerr)
        # AdamW normalises every element's gradient to ~ +-lr per step, so an element whose gradient is at the fp32-atomics
        # noise floor may move the other way: the worst case is 2*lr per step (4 steps -> 8e-3); on average far better
        diff = (flat - flat_n).abs()
        assert diff.max().item() < 2 * 4 * 1e-3 + 2e-4, (name, diff.max().item())
        assert diff.mean().item() < 5e-4, (name, diff.mean().item())
        assert (losses.cpu() - losses_n).abs().max().item() < 5e-3, name
    assert len(ran) >= 5, ran     # only nvlink_mc may be skipped (no NVLS on the box)


def _resume_worker(rank, world, port, q):
    _init(rank, world, port)
    import unet_lane_detection_b200 as U
    out = {}
    for exchange in ("nvlink_pull", "nccl"):
        out[exchange] = _resume_case(U, exchange)
    _finish(q, (rank, out))


def _resume_case(U, exchange):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g).cuda()
    y = (torch.rand(8, 1, 32, 32, generator=g) < 0.2).float().cuda()

    def fresh():
        torch.manual_seed(0)
        return U.UNet(3, 1, [64, 128]).cuda().train()

    net = fresh()
    step = U.FusedTrainStep(net, lr=1e-3, exchange=exchange, bucket_elems=20000)
    for _ in range(3):
        step.step(x, y)
    sd_opt = step.state_dict()               # collective with the sharded state: every rank
    sd_model = {k: v.clone() for k, v in net.state_dict().items()}
    step.step(x, y)
    torch.cuda.synchronize()
    want = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    # resume in a fresh model / step from the checkpoint and take the same 4th step
    net2 = fresh()
    net2.load_state_dict(sd_model)
    step2 = U.FusedTrainStep(net2, lr=1e-3, exchange=exchange, bucket_elems=20000)
    step2.load_state_dict(sd_opt, images_shape=(8, 32, 32))
    assert step2.step_count == 3
    m_full = step2._full_moments()[0].clone()        # what the resumed step holds before it moves on
    step2.step(x, y)
    torch.cuda.synchronize()
    got = torch.cat([p.detach().reshape(-1) for p in net2.parameters()])
    m_ref = torch.cat([sd_opt["state"][i]["exp_avg"].reshape(-1) for i in range(len(sd_opt["state"]))]).cuda()
    # the restored moments are the saved ones, and one more step on identical state gives the same parameters up to the
    # atomics noise (2*lr worst case per element)
    return (float((got - want).abs().max()), float((got - want).abs().mean()), float(m_ref.abs().max()),
            float((m_full - m_ref).abs().max()))


def test_two_gpu_optimizer_state_save_and_resume():
    """load_state_dict must map the global ranges onto each rank's owned parts of the sharded moments."""
    _need_two()
    res = _spawn(_resume_worker, (_port(61),), timeout=300)
    for _, out in res:
        for exchange, r in out.items():
            assert r[0] < 2e-3 + 2e-4 and r[1] < 2e-4, (exchange, r)
            assert r[2] > 0 and r[3] == 0.0, (exchange, r)


def _fit_worker(rank, world, port, q, save_dir):
    _init(rank, world, port)
    import unet_lane_detection_b200 as U
    torch.manual_seed(0)
    net = U.UNet(3, 1, [64, 128]).cuda()
    g = torch.Generator().manual_seed(10 + rank)      # every rank its own data shard
    data = [(torch.randn(8, 3, 32, 32, generator=g), (torch.rand(8, 1, 32, 32, generator=g) < 0.2).float()) for _ in range(3)]
    cfg = {"epochs": 4, "learning_rate": 1e-3, "weight_decay": 1e-4, "patience": 2, "save_dir": save_dir, "seed": 1}
    U.fit(net, data, data[:2], cfg)
    torch.cuda.synchronize()
    hist = net.b200_history
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    _finish(q, (rank, [h["val_dice"] for h in hist], float((others[0] - others[1]).abs().max()),
                os.path.exists(os.path.join(save_dir, "best_model.pth"))))


def test_two_gpu_fit_does_not_deadlock(tmp_path):
    """fit() on two ranks - the optimizer-state gather is called on every rank, the validation metrics are
    all-reduced (same best / stop decision everywhere), BatchNorm buffers come from rank 0."""
    _need_two()
    res = _spawn(_fit_worker, (_port(83), str(tmp_path)), timeout=240)
    assert res[0][1] == res[1][1] and len(res[0][1]) >= 1       # identical validation history on both ranks
    assert res[0][2] < 1e-2
    assert res[0][3]
    ck = torch.load(os.path.join(str(tmp_path), "best_model.pth"), map_location="cpu", weights_only=True)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "best_dice"}
    n_params = sum(1 for k in ck["model_state_dict"] if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert len(ck["optimizer_state_dict"]["state"]) == n_params


def test_two_executors_on_two_devices_in_one_process(tmp_path):
    """library state is per device - a container on cuda:1 after a plan on cuda:0 must set its own kernel
    attributes and use its own SM count; both give the same answer."""
    _need_two()
    sys.path.insert(0, ROOT)
    import unet_lane_detection_b200 as U
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    ref = O.UNetOracle(3, 1, [64, 128, 256, 512]).eval()
    O.randomize_bn_(ref, seed=1)
    O.scale_head_(ref, 40.0)
    path = tmp_path / "m.pth"
    torch.save(ref.state_dict(), path)
    frame = np.random.default_rng(0).integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)
    box0 = U.B200_model_container(str(path), device_id=0)
    out0 = box0.run([frame])[0]
    box1 = U.B200_model_container(str(path), device_id=1)
    out1 = box1.run([frame])[0]
    out0b = box0.run([frame])[0]
    assert np.array_equal(out0, out0b)
    assert np.abs(out0 - out1).max() <= 1e-6
    big = np.random.default_rng(1).integers(0, 256, (24, 224, 224, 3), dtype=np.uint8)     # non-graph path on both
    assert np.abs(box1.run([big])[0] - box0.run([big])[0]).max() <= 1e-6
    box0.release()
    box1.release()
