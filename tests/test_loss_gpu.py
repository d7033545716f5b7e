"""plc_combined_loss (CUDA) against the reference's CombinedLoss: committed goldens from the unmodified reference
(tests/golden/loss_*.npz, generator_*.npz) and, at training-size grids, the oracle restatement run on the same inputs.

Tolerance: fp32 reductions in a different summation order -> 2e-5 relative on the loss terms; the gradient is a sum of
+-(lambda / count) constants -> 1e-6 relative.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu
TERMS = ("point", "conserve", "smooth", "temporal")


def _close(a, b, rel=2e-5):
    return abs(float(a) - float(b)) <= rel * max(1.0, abs(float(b)))


@pytest.mark.parametrize("path", golden_files("loss_"), ids=os.path.basename)
def test_combined_loss_matches_reference_golden(path, cuda_device):
    import plconv
    g = load_golden(path)
    dev = cuda_device
    pred = torch.from_numpy(g["pred"]).to(dev).requires_grad_(True)
    mod = plconv.CombinedLoss(*[float(v) for v in g["lambdas"]], use_weighted_loss=bool(g["weighted"]),
                              weight_strategy=str(g["strategy"]))
    total, parts = mod(pred, torch.from_numpy(g["lr"]).to(dev), torch.from_numpy(g["coords"]).to(dev),
                       torch.from_numpy(g["obs"]).to(dev), scale_factor=int(g["scale"]))
    total.backward()
    for k in TERMS:
        assert _close(parts[k], g[k]), (k, float(parts[k]), float(g[k]))
    assert _close(total, g["total"])
    ref = torch.from_numpy(g["dpred"]).to(dev)
    assert torch.allclose(pred.grad, ref, rtol=1e-6, atol=1e-9), float((pred.grad - ref).abs().max())


@pytest.mark.parametrize("path", golden_files("generator_"), ids=os.path.basename)
def test_combined_loss_on_generator_golden(path, cuda_device):
    """Loss terms the reference computed on its own Generator output (CPU coordinates / observations accepted)."""
    import plconv
    g = load_golden(path)
    total, parts = plconv.CombinedLoss()(torch.from_numpy(g["pred"]).to(cuda_device),
                                         torch.from_numpy(g["rain"]).to(cuda_device), torch.from_numpy(g["s_coords"]),
                                         torch.from_numpy(g["s_vals"]), scale_factor=int(g["scale"]))
    for k in TERMS:
        assert _close(parts[k], g["loss_" + k]), k
    assert _close(total, g["loss_total"])


@pytest.mark.parametrize("B,T,H,W,s,n_st", [(4, 5, 30, 24, 4, 60), (2, 3, 128, 128, 1, 300), (1, 1, 16, 16, 2, 5)])
def test_combined_loss_vs_oracle_large(B, T, H, W, s, n_st, cuda_device):
    import plconv
    from oracle import loss_oracle as L
    gen = torch.Generator().manual_seed(B * 100 + H)
    pred = (torch.rand(B, T, 1, H * s, W * s, generator=gen) * 10).to(cuda_device)
    lr = (torch.rand(B, T, 1, H, W, generator=gen) * 10).to(cuda_device)
    coords = torch.stack([torch.randint(-1, H + 1, (n_st,), generator=gen),
                          torch.randint(-1, W + 1, (n_st,), generator=gen)], 1).to(cuda_device)
    obs = (torch.rand(T, n_st, generator=gen) * 40).to(cuda_device)
    obs[torch.rand(T, n_st, generator=gen).to(cuda_device) < 0.2] = float("nan")
    p1 = pred.clone().requires_grad_(True)
    p2 = pred.clone().requires_grad_(True)
    total, parts = plconv.CombinedLoss()(p1, lr, coords, obs, scale_factor=s)
    want, wparts = L.combined_loss(p2, lr, coords, obs, scale_factor=s)
    for k in TERMS:
        a, b = float(parts[k]), float(wparts[k])
        assert (np.isnan(a) and np.isnan(b)) or _close(a, b), (k, a, b)    # T == 1: temporal mean of nothing is NaN
    if T > 1:
        total.backward()
        want.backward()
        assert torch.allclose(p1.grad, p2.grad, rtol=1e-5, atol=1e-10), float((p1.grad - p2.grad).abs().max())


def test_combined_loss_edge_cases(cuda_device):
    import plconv
    dev = cuda_device
    pred = torch.rand(2, 3, 1, 8, 8, device=dev, requires_grad=True)
    lr = torch.rand(2, 3, 1, 4, 4, device=dev)
    mod = plconv.CombinedLoss()
    # no stations / no observations -> point term 0 (combined_loss.py:83-84)
    for coords, obs in ((torch.zeros(0, 2, dtype=torch.long), torch.zeros(3, 0)), (torch.zeros(4, 2, dtype=torch.long), None)):
        total, parts = mod(pred, lr, coords, obs, scale_factor=2)
        assert float(parts["point"]) == 0.0 and torch.isfinite(total)
    # every observation missing -> 0 (combined_loss.py:127-128), gradient still defined
    total, parts = mod(pred, lr, torch.zeros(4, 2, dtype=torch.long), torch.full((3, 4), float("nan")), scale_factor=2)
    total.backward()
    assert float(parts["point"]) == 0.0 and torch.isfinite(pred.grad).all()
    # gradient scales with the upstream gradient
    g1 = pred.grad.clone()
    pred.grad = None
    total, _ = mod(pred, lr, torch.zeros(4, 2, dtype=torch.long), torch.full((3, 4), float("nan")), scale_factor=2)
    (3.0 * total).backward()
    assert torch.allclose(pred.grad, 3.0 * g1)
    # loud errors: CPU tensor, non-integer ratio
    with pytest.raises(RuntimeError, match="CUDA"):
        mod(pred.detach().cpu(), lr.cpu(), None, None)
    with pytest.raises(RuntimeError, match="integer"):
        mod(torch.rand(1, 1, 1, 9, 9, device=dev), torch.rand(1, 1, 1, 4, 4, device=dev), None, None)
