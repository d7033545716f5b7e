"""Every ABI call must run on the device (and that device's current stream) its tensors live on, not on the process's
current device (`Trainer(device='cuda:1')` without `torch.cuda.set_device`).  Needs two GPUs; skipped on a
single-GPU box."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def second_device():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    assert torch.cuda.current_device() == 0
    return torch.device("cuda:1")


def test_cell_step_and_bptt_on_the_non_current_device(second_device):
    import plconv
    from plconv import functional as F
    g = torch.Generator().manual_seed(3)
    B, cin, ch, H, W, k = 2, 64, 64, 24, 20, 3
    w = (torch.rand(4 * ch, cin + ch, k, k, generator=g) - 0.5) * 0.1
    b = torch.randn(4 * ch, generator=g) * 0.2
    x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16)
    h = (torch.randn(B, H, W, ch, generator=g) * 0.5).to(torch.bfloat16)
    c = torch.randn(B, H, W, ch, generator=g)
    dh = torch.randn(B, H, W, ch, generator=g).to(torch.bfloat16)
    dc = torch.randn(B, H, W, ch, generator=g)

    def run(dev):
        pw = F.pack_weights(w.to(dev), b.to(dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
        xd, hd, cd = x.to(dev), h.to(dev), c.to(dev)
        saved = torch.empty(F.saved_gates_bytes(B, H, W, pw), dtype=torch.uint8, device=dev)
        h2, c2 = F.cell_forward(xd, hd, cd, pw, saved=saved)
        img = F.wgrad_accumulator(B, H, W, pw, dev)
        db = torch.zeros(4 * ch, device=dev)
        dx, dhp, dcp = F.cell_backward_acc(xd, hd, cd, pw, dh.to(dev), None, dc.to(dev), img, db, saved=saved)
        torch.cuda.synchronize(dev)
        return [t.cpu() for t in (h2, c2, dx, dhp, dcp, db)]

    a = run(torch.device("cuda:0"))
    bb = run(second_device)                                   # current device stays cuda:0 throughout
    assert torch.cuda.current_device() == 0
    for i, (p, q) in enumerate(zip(a, bb)):
        if i < 5:
            assert torch.equal(p, q), i                      # same kernels, same tiles: bit-identical
        else:
            assert torch.allclose(p, q, rtol=1e-4, atol=1e-4)   # db: atomics order


def test_generator_training_step_on_the_non_current_device(second_device):
    import plconv
    torch.manual_seed(9)
    model = plconv.NowcastGenerator(1, [64, 64], 3, 3, 3, "bf16").to(second_device)
    frames = torch.relu(torch.randn(2, 3, 1, 32, 32, device=second_device) + 0.3)
    target = torch.relu(torch.randn(2, 3, 1, 32, 32, device=second_device) + 0.3)
    loss = (model(frames) - target).abs().mean()
    loss.backward()
    torch.cuda.synchronize(second_device)
    assert torch.cuda.current_device() == 0
    assert torch.isfinite(loss) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    ref = plconv.NowcastGenerator(1, [64, 64], 3, 3, 3, "bf16").to("cuda:0")
    ref.load_state_dict(model.state_dict())
    loss0 = (ref(frames.to("cuda:0")) - target.to("cuda:0")).abs().mean()
    assert abs(float(loss0.detach()) - float(loss.detach())) < 1e-6
