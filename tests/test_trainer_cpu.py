"""Host logic of plconv.trainer on CPU (stub model: the CUDA kernels are not involved): early stopping against the
live reference class, epoch loop / scheduler / best-model checkpoint format (trainer.py:402-418) / resume, the
sync-free station RMSE against the reference formula, and the world-size-2 gloo path (collective-safe NaN-skip)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

import plconv
from plconv.trainer import EarlyStopping, Trainer, TrainerConfig, station_rmse

REF = os.environ.get("PLC_REFERENCE", "/root/reference")


class _StubGen(torch.nn.Module):
    """rain [B,T,1,H,W] -> [B,T,1,2H,2W]; two top-level children so there are two gradient buckets."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(1, 4, 3, padding=1)
        self.b = torch.nn.Conv2d(4, 1, 3, padding=1)

    def forward(self, rain, dem, lu):
        B, T = rain.shape[:2]
        x = torch.nn.functional.interpolate(rain.flatten(0, 1), scale_factor=2.0, mode="nearest")
        return self.b(torch.relu(self.a(x))).view(B, T, 1, *x.shape[-2:])


class _StubLoss(torch.nn.Module):
    def forward(self, pred, lr, s_coords, s_values, scale_factor=1.0):
        pooled = torch.nn.functional.avg_pool2d(pred.flatten(0, 1), 2).view_as(lr)
        cons = (pooled - lr).abs().mean()
        z = cons.detach() * 0
        return cons, {"point": z, "conserve": cons.detach(), "smooth": z, "temporal": z}


def _batches(n, seed=0, B=2, nan_at=None):
    g = torch.Generator().manual_seed(seed)
    coords = torch.tensor([[0, 0], [3, 4], [9, 9]])
    out = []
    for i in range(n):
        rain = torch.rand(B, 2, 1, 6, 8, generator=g)
        if nan_at == i:
            rain[0, 0, 0, 0, 0] = float("nan")
        out.append((rain, torch.zeros(B, 1, 12, 16), torch.zeros(B, 0, 12, 16), coords,
                    torch.rand(B, 2, 3, generator=g)))
    return out


def _trainer(tmp=None, **kw):
    torch.manual_seed(1)
    cfg = TrainerConfig(scale_factor=2, output_dir=tmp, **kw)
    return Trainer(cfg, device="cpu", model=_StubGen(), loss_module=_StubLoss())


def test_early_stopping_matches_reference_class():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from src.utils.early_stopping import EarlyStopping as RefES
    scores = [5.0, 4.0, 4.0, 3.99, 4.5, 3.0, 3.2, 3.1, 3.05, 2.0]
    for patience, delta in ((2, 0.0), (3, 0.05), (1, 0.0)):
        a, b = EarlyStopping(patience, delta), RefES(patience=patience, min_delta=delta, verbose=False)
        for e, s in enumerate(scores):
            assert a(s, e) == b(s, e)
            assert (a.counter, a.best_score, a.best_epoch, a.early_stop) == (b.counter, b.best_score, b.best_epoch,
                                                                             b.early_stop)


def test_station_rmse_matches_reference_formula():
    torch.manual_seed(0)
    fake = torch.rand(2, 3, 1, 12, 16)
    coords = torch.tensor([[0, 0], [5, 7], [6, 2], [2, 7]])          # [6,2]*2 -> row 12: off the 12-row grid
    obs = torch.rand(2, 3, 4)
    obs[0, 1, 1] = float("nan")
    got = station_rmse(fake, coords, obs, 2.0)
    sc = ((coords.float() + 0.5) * 2.0 - 0.5).long()                 # trainer.py:236-262
    ok = (sc[:, 0] < 12) & (sc[:, 1] < 16)
    at = fake[:, :, 0][:, :, sc[ok, 0], sc[ok, 1]]
    tv = obs[:, :, ok]
    m = ~torch.isnan(tv)
    want = torch.sqrt(torch.nn.functional.mse_loss(at[m], tv[m]))
    assert torch.allclose(got, want, atol=1e-7)
    assert float(station_rmse(fake, coords, torch.full((2, 3, 4), float("nan")), 2.0)) == 0.0


def test_fit_checkpoint_format_and_resume(tmp_path):
    tr = _trainer(str(tmp_path), epochs=3)
    hist = tr.fit(_batches(4), _batches(2, seed=9))
    assert hist["epoch"] == [0, 1, 2] and len(hist["total_loss"]) == 3
    assert hist["total_loss"][2] < hist["total_loss"][0]
    ck = torch.load(tmp_path / "best_model.pth", weights_only=False)
    for key in ("epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "rmse", "history"):
        assert key in ck                                            # trainer.py:410-417
    assert set(ck["model_state_dict"]) == set(tr.model.state_dict())
    # resume: a fresh trainer continues at the next epoch with the saved moments / scheduler / history
    tr2 = _trainer(str(tmp_path), epochs=5)
    tr2.load_checkpoint(str(tmp_path / "best_model.pth"))
    assert tr2.start_epoch == ck["epoch"] + 1 and tr2.best_rmse == pytest.approx(ck["best_rmse"])
    for k, v in ck["model_state_dict"].items():
        assert torch.equal(tr2.model.state_dict()[k], v)
    assert tr2.optimizer.state_dict()["state"][0]["step"] == ck["optimizer_state_dict"]["state"][0]["step"]
    hist2 = tr2.fit(_batches(4), _batches(2, seed=9))
    assert hist2["epoch"][-1] == 4 and len(hist2["epoch"]) == len(ck["history"]["epoch"]) + (5 - tr.best_epoch - 1)
    # a checkpoint with only the reference's six keys loads too
    ref_style = {k: ck[k] for k in ("epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict",
                                    "rmse", "history")}
    torch.save(ref_style, tmp_path / "ref_style.pth")
    tr3 = _trainer(None, epochs=1)
    tr3.load_checkpoint(str(tmp_path / "ref_style.pth"))
    assert tr3.best_rmse == pytest.approx(ck["rmse"])


def test_scheduler_and_early_stopping_drive_the_loop():
    tr = _trainer(None, epochs=50, early_stopping_patience=2, scheduler_patience=0, scheduler_factor=0.5,
                  learning_rate=0.0)                                # lr 0: the metric never improves
    hist = tr.fit(_batches(2))
    assert len(hist["epoch"]) == 3 and tr.early_stopping.early_stop   # best at epoch 0, then 2 stale epochs


def test_nan_batch_is_skipped_like_the_reference():
    tr = _trainer(None, epochs=1)
    before = [p.detach().clone() for p in tr.model.parameters()]
    assert tr.train_step(tuple(_batches(1, nan_at=0)[0])) is None and tr.skipped == 1   # trainer.py:306-308
    assert all(torch.equal(a, b) for a, b in zip(before, tr.model.parameters()))
    assert tr.train_step(tuple(_batches(1)[0])) is not None


def _ddp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                               # different init per rank: must be broadcast
        tr = Trainer(TrainerConfig(scale_factor=2, epochs=2), device="cpu", model=_StubGen(), loss_module=_StubLoss())
        # rank 1 sees a NaN batch at step 1: BOTH ranks must skip it (collective-safe), then keep training
        data = _batches(3, seed=rank, B=1, nan_at=1 if rank == 1 else None)
        tr.fit(data)
        flat = torch.cat([p.detach().flatten() for p in tr.model.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        torch.distributed.all_gather(gathered, flat)
        if rank == 0:
            ret["same"] = bool(torch.equal(gathered[0], gathered[1]))
            ret["skipped"] = tr.skipped
            ret["loss"] = tr.history["total_loss"]
    finally:
        torch.distributed.destroy_process_group()


def test_trainer_world_size_2_gloo():
    port = 29650 + os.getpid() % 200
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ddp_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["same"], "ranks diverged"
        assert ret["skipped"] == 2                                  # one skipped step in each of the 2 epochs
        assert len(ret["loss"]) == 2


def test_single_process_matches_two_rank_average():
    """Gradient averaging over 2 ranks with batch 1 each == one process with the batch of 2 (mean losses)."""
    torch.manual_seed(1)
    tr = _trainer(None, epochs=1)
    b0, b1 = _batches(1, seed=0, B=1)[0], _batches(1, seed=1, B=1)[0]
    both = tuple(torch.cat([x, y]) if x.dim() > 2 else x for x, y in zip(b0, b1))
    tr.train_step(both)
    ref = [p.detach().clone() for p in tr.model.parameters()]
    # emulate the two ranks by hand: mean of per-rank gradients
    tr2 = _trainer(None, epochs=1)
    grads = []
    for b in (b0, b1):
        loss, _ = tr2.loss_module(tr2.model(*b[:3]), b[0], b[3], b[4], 2.0)
        grads.append(torch.autograd.grad(loss, list(tr2.model.parameters())))
    tr2.reducer.zero_grad()
    for p, g0, g1 in zip(tr2.model.parameters(), *grads):
        p.grad.copy_((g0 + g1) / 2)
    torch.nn.utils.clip_grad_norm_(tr2.trainable, 0.5)
    tr2.optimizer.step()
    for a, b in zip(ref, tr2.model.parameters()):
        assert torch.allclose(a, b, atol=1e-7)


def test_fused_optimizer_step_invalidates_packed_weight_caches():
    """torch's fused optimizers update parameters without bumping their autograd version counter, which the
    packed-weight caches are keyed on; the global post-step hook must advance the cache generation instead."""
    from plconv import _lib
    p = torch.nn.Parameter(torch.randn(4))
    opt = torch.optim.Adam([p], lr=1e-3, fused=True)
    p.grad = torch.randn(4)
    g0 = _lib.weight_generation()
    opt.step()
    assert _lib.weight_generation() > g0
    g1 = _lib.weight_generation()
    plconv.invalidate_packed_weights()
    assert _lib.weight_generation() == g1 + 1


class _LazyStubGen(_StubGen):
    """Like the reference Generator: one block only exists after ``materialize`` (generator.py:129-130)."""

    def __init__(self):
        super().__init__()
        self.up = None

    def materialize(self, scale_factor, device=None):
        if self.up is None:
            torch.manual_seed(77)
            self.up = torch.nn.Conv2d(1, 1, 3, padding=1)
        return 1

    def forward(self, rain, dem, lu):
        out = super().forward(rain, dem, lu)
        B, T = out.shape[:2]
        return self.up(out.flatten(0, 1)).view_as(out)


def _quirk_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(5)
        tr = Trainer(TrainerConfig(scale_factor=2, epochs=1, optimizer_sees_upsample=False), device="cpu",
                     model=_LazyStubGen(), loss_module=_StubLoss())
        up0 = [p.detach().clone() for p in tr.model.up.parameters()]
        data = _batches(3, seed=10 + rank, B=1)
        for b in data:
            tr.train_step(tuple(b))
        flat = torch.cat([p.detach().flatten() for p in tr.model.parameters()])
        fg = tr.frozen_reducer.buckets[0]["flat"].clone()
        both, bothg = [torch.zeros_like(flat) for _ in range(world)], [torch.zeros_like(fg) for _ in range(world)]
        torch.distributed.all_gather(both, flat)
        torch.distributed.all_gather(bothg, fg)
        if rank == 0:
            ret["params_same"] = bool(torch.equal(both[0], both[1]))
            ret["frozen_grads_same"] = bool(torch.allclose(bothg[0], bothg[1], rtol=1e-6, atol=1e-9))
            ret["frozen_untouched"] = all(torch.equal(a, b) for a, b in zip(up0, tr.model.up.parameters()))
            ret["frozen_grad_norm"] = float(fg.norm())
    finally:
        torch.distributed.destroy_process_group()


def test_reference_quirk_mode_under_data_parallelism():
    """optimizer_sees_upsample=False with 2 ranks: the never-optimised block stays at its initial values, its
    never-zeroed gradient buffer is averaged consistently (so every rank computes the same clip coefficient), and the
    trained parameters stay identical across ranks."""
    port = 29850 + os.getpid() % 100
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_quirk_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["params_same"] and ret["frozen_grads_same"] and ret["frozen_untouched"]
        assert ret["frozen_grad_norm"] > 0


def test_reference_quirk_mode_single_process_accumulates_and_clips_frozen_grads():
    """Same quirk, one process, against a hand-rolled replay of trainer.py:290-315 with a late-created block."""
    torch.manual_seed(5)
    tr = Trainer(TrainerConfig(scale_factor=2, epochs=1, optimizer_sees_upsample=False), device="cpu",
                 model=_LazyStubGen(), loss_module=_StubLoss())
    torch.manual_seed(5)
    ref = _LazyStubGen()
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4)          # built BEFORE the block exists, as the reference does
    ref.materialize(2)
    ref.load_state_dict(tr.model.state_dict())
    loss_mod = _StubLoss()
    for b in _batches(3, seed=3):
        tr.train_step(tuple(b))
        opt.zero_grad()
        loss, _ = loss_mod(ref(*b[:3]), b[0], b[3], b[4], 2.0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt.step()
    for (n, a), (_, b_) in zip(tr.model.named_parameters(), ref.named_parameters()):
        assert torch.allclose(a, b_, atol=1e-7), n
    for a, b_ in zip(tr.model.up.parameters(), ref.up.parameters()):
        assert torch.allclose(a.grad, b_.grad, rtol=1e-5, atol=1e-8)      # accumulated and rescaled by every clip


def test_reference_written_checkpoint_resumes_with_real_generator(tmp_path):
    """A checkpoint as the REFERENCE writes it (trainer.py:402-418): Adam built before ``upsample_blocks`` exist, so
    its param_groups hold only the leading parameters.  Resuming with the default optimizer_sees_upsample=True must
    restore those moments, give the lazy blocks fresh ones, and seed early stopping from the saved rmse."""
    torch.manual_seed(3)
    kw = dict(in_channels=1, dem_channels=1, lu_channels=2, hidden_dims=[16, 32], scale_factor=4)
    ref_like = plconv.Generator(mode="fp32", **kw)
    opt = torch.optim.Adam(ref_like.parameters(), lr=5e-4)          # trainer.py:153-158: BEFORE the first forward
    n_before = len(opt.param_groups[0]["params"])
    ref_like.materialize(4)                                          # what the first forward does (generator.py:129)
    for p in opt.param_groups[0]["params"]:
        p.grad = torch.randn_like(p)
    opt.step()
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.7, patience=10)
    path = tmp_path / "best_model.pth"
    torch.save({"epoch": 6, "model_state_dict": ref_like.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "scheduler_state_dict": sched.state_dict(), "rmse": 1.25, "history": {"epoch": list(range(7))}}, path)

    cfg = TrainerConfig(hidden_dims=(16, 32), lu_channels=2, scale_factor=4, mode="fp32", early_stopping_patience=5)
    tr = Trainer(cfg, device="cpu", model=plconv.Generator(mode="fp32", **kw), loss_module=_StubLoss())
    n_all = len(tr.optimizer.param_groups[0]["params"])
    assert n_all == n_before + 4                                     # two x2 blocks: weight + bias each
    tr.load_checkpoint(str(path))
    state = tr.optimizer.state_dict()["state"]
    assert sorted(state) == list(range(n_before))                    # restored moments for the leading parameters only
    want = opt.state_dict()["state"]
    for i in range(n_before):
        assert torch.equal(state[i]["exp_avg"], want[i]["exp_avg"])
    for a, b in zip(tr.optimizer.param_groups[0]["params"], tr.model.parameters()):
        assert a is b
    assert tr.start_epoch == 7 and tr.best_rmse == pytest.approx(1.25)
    assert tr.early_stopping.best_score == pytest.approx(1.25)       # the first resumed epoch must beat it to save
    assert tr.early_stopping(1.3, 7) is False
    # and the step after the resume runs (fresh moments are created lazily for the upsample parameters)
    for p in tr.optimizer.param_groups[0]["params"]:
        p.grad = torch.zeros_like(p)
    tr.optimizer.step()
    assert len(tr.optimizer.state_dict()["state"]) == n_all

    quirk = Trainer(TrainerConfig(hidden_dims=(16, 32), lu_channels=2, scale_factor=4, mode="fp32",
                                  optimizer_sees_upsample=False), device="cpu",
                    model=plconv.Generator(mode="fp32", **kw), loss_module=_StubLoss())
    quirk.load_checkpoint(str(path))                                 # same parameter count as the reference: direct load
    full = tmp_path / "full.pth"
    tr.save_checkpoint(str(full), 7, 1.0)
    with pytest.raises(RuntimeError, match="optimizer_sees_upsample"):
        quirk.load_checkpoint(str(full))
