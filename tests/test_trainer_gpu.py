"""plconv.Trainer on the GPU against the reference's optimisation loop (tests/golden/train_*.npz: unmodified
Generator + CombinedLoss + Adam + clip replayed on CPU by make_golden.py), plus the properties the B200 loop adds:
device-side NaN-skip, no host synchronisation inside a step, checkpoint round trip through the real Generator."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _trainer_from_golden(g, mode, dev, quirk=True):
    import plconv
    cfg = plconv.TrainerConfig(hidden_dims=[int(v) for v in g["hidden_dims"]], lu_channels=int(g["lu_ch"]),
                               scale_factor=int(g["scale"]), mode=mode, optimizer_sees_upsample=not quirk)
    tr = plconv.Trainer(cfg, device=dev)
    # initial state: sd0 (before training) + the never-updated upsample blocks from the final state
    sd = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd1.")}
    sd.update({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd0.")})
    tr.model.load_state_dict(sd)
    return tr


@pytest.mark.parametrize("path", golden_files("train_"), ids=os.path.basename)
def test_training_steps_replay_reference_loop_fp32(path, cuda_device):
    """Quirk mode (optimizer built before upsample_blocks exist), fp32: per-step loss terms and station RMSE within
    1e-3 rel. of the reference, parameters after 4 Adam steps within 5e-5 abs (lr = 5e-4; Adam's first steps move
    every weight by ~lr*sign(grad), so only weights whose gradient is ~0 can differ at all), untouched upsample
    blocks bit-identical."""
    g = load_golden(path)
    dev = cuda_device
    tr = _trainer_from_golden(g, "fp32", dev)
    coords = torch.from_numpy(g["coords"]).to(dev)
    for i in range(int(g["steps"])):
        batch = tuple(torch.from_numpy(g[f"{k}{i}"]).to(dev) for k in ("rain", "dem", "lu")) + \
            (coords, torch.from_numpy(g[f"obs{i}"]).to(dev))
        row = tr.train_step(batch).cpu().numpy()
        want = np.concatenate([g[f"loss{i}"], g[f"rmse{i}"].reshape(1)])
        assert np.allclose(row[:6], want, rtol=1e-3, atol=1e-5), (i, row[:6], want)
        assert row[6] == 1.0
    worst = {}
    for k, v in tr.model.state_dict().items():
        ref = torch.from_numpy(g["sd1." + k]).to(dev)
        if k.startswith("upsample_blocks"):
            assert torch.equal(v, ref), k                               # never in the optimizer (generator.py:129-130)
        else:
            worst[k] = float((v - ref).abs().max())
            moved = float((ref - torch.from_numpy(g["sd0." + k]).to(dev)).abs().max())
            assert moved > 1e-4, k                                      # the reference did train this tensor
    bad = {k: e for k, e in worst.items() if e > 5e-5}
    assert not bad, bad


@pytest.mark.parametrize("path", golden_files("train_"), ids=os.path.basename)
def test_training_steps_bf16_track_reference_losses(path, cuda_device):
    """bf16 tensor-core mode on the same batches: losses within 2 % of the reference's fp32 loop at every step."""
    g = load_golden(path)
    tr = _trainer_from_golden(g, "bf16", cuda_device)
    coords = torch.from_numpy(g["coords"]).to(cuda_device)
    for i in range(int(g["steps"])):
        batch = tuple(torch.from_numpy(g[f"{k}{i}"]).to(cuda_device) for k in ("rain", "dem", "lu")) + \
            (coords, torch.from_numpy(g[f"obs{i}"]).to(cuda_device))
        row = tr.train_step(batch).cpu().numpy()
        assert np.allclose(row[:5], g[f"loss{i}"], rtol=2e-2, atol=1e-4), (i, row[:5], g[f"loss{i}"])


def _small_trainer(dev, tmp=None, **kw):
    import plconv
    torch.manual_seed(0)
    cfg = plconv.TrainerConfig(hidden_dims=[16, 16], lu_channels=3, scale_factor=2, output_dir=tmp, **kw)
    data = plconv.trainer.SyntheticRainBatches(4, B=2, T=3, H=8, W=10, scale=2, lu_channels=3, n_stations=6, seed=3)
    return plconv.Trainer(cfg, device=dev), data


def test_step_has_no_host_synchronisation_and_loss_decreases(cuda_device):
    tr, data = _small_trainer(cuda_device, epochs=6)
    tr.train_epoch(data)                                              # warm-up: lazy packing, cuDNN/cuBLAS handles
    batches = list(plconv_prefetch(tr, data))
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for b in batches:
            tr.train_step(b)                                          # raises on any synchronising CUDA call
    finally:
        torch.cuda.set_sync_debug_mode("default")
    hist = tr.fit(data)
    assert hist["total_loss"][-1] < hist["total_loss"][0]
    assert tr.skipped == 0


def plconv_prefetch(tr, data):
    from plconv.trainer import DevicePrefetcher
    return DevicePrefetcher(data, tr.device)


def test_nan_batch_skipped_on_device(cuda_device):
    tr, data = _small_trainer(cuda_device, epochs=1)
    good = next(iter(plconv_prefetch(tr, data)))
    tr.train_step(good)
    before = [p.detach().clone() for p in tr.model.parameters()]
    steps_before = [s["step"].clone() for s in tr.optimizer.state.values()]
    bad = (good[0].clone(),) + tuple(good[1:])
    bad[0][0, 0, 0, 0, 0] = float("nan")
    row = tr.train_step(bad)
    assert float(row[6]) == 0.0 and torch.isfinite(row).all()         # not counted, nothing poisoned
    assert all(torch.equal(a, b) for a, b in zip(before, tr.model.parameters()))
    assert all(torch.equal(a, s["step"]) for a, s in zip(steps_before, tr.optimizer.state.values()))
    row = tr.train_step(good)
    assert float(row[6]) == 1.0 and all(torch.isfinite(p).all() for p in tr.model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.model.parameters()))


def test_checkpoint_round_trip_real_generator(tmp_path, cuda_device):
    tr, data = _small_trainer(cuda_device, str(tmp_path), epochs=2)
    tr.fit(data)
    path = tmp_path / "best_model.pth"
    ck = torch.load(path, weights_only=False)
    assert any(k.startswith("upsample_blocks.") for k in ck["model_state_dict"])
    assert {"cell1.conv.weight", "cell2.conv.bias"} <= set(ck["model_state_dict"])
    tr2, _ = _small_trainer(cuda_device, None, epochs=3)
    tr2.load_checkpoint(str(path))
    batch = next(iter(plconv_prefetch(tr2, data)))
    tr.model.load_state_dict(ck["model_state_dict"])
    with torch.no_grad():
        a, b = tr.model(*batch[:3]), tr2.model(*batch[:3])
    assert torch.equal(a, b)                                          # packed-weight caches rebuilt after the load
    assert tr2.start_epoch == ck["epoch"] + 1
    tr2.fit(data)
    assert tr2.history["epoch"][-1] == 2
