"""Discriminator / GAN step on the GPU vs the repo-defined eager spec (oracle/gan_oracle.py).

PARITY VS REPO SPEC, NOT VS THE REFERENCE: the reference has no discriminator, no Conv3d and no adversarial loss
(SURVEY.md section 0).  Everything goes through the C ABI (plc_convnd_* -> conv_igemm_tc_kernel with strided / 5-D
tensor maps).  Tolerances are relative to the tensor's max |value| (global-max norm) on bf16-rounded operands:
forward 1e-2, gradients 2e-2.
"""
import pytest
import torch
import torch.nn.functional as TF

from test_cell_gpu import rel_err, report

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


# (B, T, H, W, Cin, Cout, kt, k, stride_t, stride, act)
ND_CASES = [
    (3, 1, 32, 32, 1, 32, 1, 3, 1, 2, 2),       # conv1: strided 2-D, 1 (-> 8 padded) input channel, LeakyReLU
    (2, 1, 19, 13, 8, 16, 1, 3, 1, 2, 1),       # ragged odd grid, ReLU
    (2, 1, 16, 24, 128, 1, 1, 3, 1, 1, 0),      # score conv: stride 1, no activation, 1 (-> 8) output channel
    (2, 6, 16, 16, 32, 64, 3, 3, 1, 2, 2),      # conv2: 3-D, time stride 1, space stride 2
    (2, 6, 16, 16, 64, 128, 3, 3, 2, 2, 2),     # conv3: 3-D, strides (2, 2, 2)
    (1, 5, 11, 9, 16, 24, 3, 3, 2, 2, 2),       # 3-D, odd T / H / W (ragged everything)
    (1, 4, 12, 12, 16, 16, 3, 3, 1, 1, 0),      # 3-D, all strides 1
    (1, 3, 20, 20, 8, 32, 1, 5, 1, 2, 2),       # 5x5 strided 2-D over 3 frames
    (2, 1, 64, 64, 64, 128, 1, 3, 1, 2, 2),     # enough tiles for the CTA-pair path when forced
]


def _run_nd(case, dev):
    from plconv import functional as PF
    B, T, H, W, cin, cout, kt, k, st, s, act = case
    torch.manual_seed(7)
    if kt > 1 or st > 1:
        conv = torch.nn.Conv3d(cin, cout, (kt, k, k), stride=(st, s, s), padding=(kt // 2, k // 2, k // 2)).to(dev)
    else:
        conv = torch.nn.Conv2d(cin, cout, k, stride=s, padding=k // 2).to(dev)
    cp = PF.ConvNdParams(conv, act=act, slope=0.2)
    x = torch.randn(B, T, H, W, cin, device=dev)
    xw = TF.pad(x, (0, cp.cin_p - cin)).to(torch.bfloat16).contiguous().requires_grad_()
    if cp.is3d:
        y = PF.convnd(xw, cp)                                              # [B,To,Ho,Wo,cout_p]
    else:
        y = PF.convnd(xw.view(B * T, 1, H, W, cp.cin_p), cp)
        y = y.view(B, T, *y.shape[2:])
    # torch restatement on the same bf16-rounded operands
    xr = xw.detach()[..., :cin].float().requires_grad_()
    wr = conv.weight.detach().to(torch.bfloat16).float().requires_grad_()
    br = conv.bias.detach().clone().requires_grad_()
    if cp.is3d:
        yr = TF.conv3d(xr.permute(0, 4, 1, 2, 3), wr, br, stride=(st, s, s), padding=(kt // 2, k // 2, k // 2))
        yr = yr.permute(0, 2, 3, 4, 1)
    else:
        yr = TF.conv2d(xr.reshape(B * T, H, W, cin).permute(0, 3, 1, 2), wr, br, stride=s, padding=k // 2)
        yr = yr.permute(0, 2, 3, 1).reshape(B, T, yr.shape[2], yr.shape[3], cout)
    if act == 1:
        yr = torch.relu(yr)
    elif act == 2:
        yr = TF.leaky_relu(yr, 0.2)
    assert tuple(y.shape[:4]) == tuple(yr.shape[:4]), (y.shape, yr.shape)
    assert rel_err(y[..., :cout], yr) < 1e-2, report("y", y[..., :cout], yr)
    if y.shape[-1] > cout:
        assert float(y.detach()[..., cout:].abs().max()) == 0.0           # padded output channels stay zero
    gy = torch.randn_like(yr)
    gyw = torch.zeros_like(y)
    gyw[..., :cout] = gy.to(torch.bfloat16)
    y.backward(gyw)
    # the activation mask is taken from the bf16 output the kernel stored; use the same sign pattern in the restatement
    (yr * gyw[..., :cout].float()).sum().backward()
    assert rel_err(xw.grad[..., :cin], xr.grad) < 2e-2, report("dx", xw.grad[..., :cin], xr.grad)
    assert rel_err(conv.weight.grad, wr.grad) < 2e-2, report("dW", conv.weight.grad, wr.grad)
    assert rel_err(conv.bias.grad, br.grad) < 2e-2, report("db", conv.bias.grad, br.grad)


@pytest.mark.parametrize("case", ND_CASES, ids=lambda c: "B%d_T%d_%dx%d_%d-%d_kt%d_k%d_st%d_s%d_act%d" % c)
def test_strided_and_3d_conv_forward_backward_vs_torch(case, cuda_device):
    _run_nd(case, cuda_device)


@pytest.mark.parametrize("cta", [1, 2])
@pytest.mark.parametrize("case", [ND_CASES[0], ND_CASES[4], ND_CASES[5], ND_CASES[8]],
                         ids=lambda c: "B%d_T%d_%dx%d_%d-%d_kt%d_k%d_st%d_s%d_act%d" % c)
def test_strided_and_3d_conv_both_cta_group_paths(case, cta, cuda_device):
    import plconv
    lib = plconv._lib.load()
    lib.plc_debug_set_cta_group(cta)
    try:
        _run_nd(case, cuda_device)
    finally:
        lib.plc_debug_set_cta_group(0)


# (N, Cf, H, W, Cout, stride, act)
FRAME_CASES = [
    (6, 1, 32, 32, 32, 2, 2),        # the discriminator's conv1
    (3, 1, 19, 13, 32, 2, 2),        # odd, ragged grid
    (2, 3, 17, 24, 16, 2, 1),        # 3 frame channels, ReLU
    (2, 2, 16, 21, 8, 1, 0),         # stride 1, no activation
    (2, 4, 12, 12, 64, 1, 2),        # 4 channels, stride 1
    (1, 1, 9, 7, 128, 2, 2),
]


@pytest.mark.parametrize("case", FRAME_CASES, ids=lambda c: "N%d_Cf%d_%dx%d_C%d_s%d_act%d" % c)
def test_frameconv_forward_backward_vs_torch(case, cuda_device):
    """plc_frameconv_fwd / _bwd (first discriminator layer straight from fp32 frames) vs torch conv2d in fp64: output
    (bf16 store: 1e-2 of max), d frames, dW, db (the activation mask taken from the stored bf16 output on both sides)."""
    from plconv import functional as PF
    n, cf, hh, ww, cout, s, act = case
    torch.manual_seed(11)
    conv = torch.nn.Conv2d(cf, cout, 3, stride=s, padding=1).to(cuda_device)
    assert PF.frameconv_supported(conv)
    x = torch.randn(n, cf, hh, ww, device=cuda_device).requires_grad_()
    y = PF.frameconv(x, conv, act=act, slope=0.2)                          # [n, ho, wo, cout] bf16
    xr = x.detach().double().cpu().requires_grad_()
    wr = conv.weight.detach().double().cpu().requires_grad_()
    br = conv.bias.detach().double().cpu().requires_grad_()
    zr = TF.conv2d(xr, wr, br, stride=s, padding=1)
    pos = (y.detach() > 0).permute(0, 3, 1, 2).cpu()
    yr = zr if act == 0 else torch.where(pos, zr, (0.2 if act == 2 else 0.0) * zr)
    assert tuple(y.shape) == (n, yr.shape[2], yr.shape[3], cout)
    assert rel_err(y.permute(0, 3, 1, 2).cpu(), yr) < 1e-2, report("y", y.permute(0, 3, 1, 2).cpu(), yr)
    gy = torch.randn_like(y)
    y.backward(gy)
    (yr * gy.float().permute(0, 3, 1, 2).cpu().double()).sum().backward()
    assert rel_err(x.grad.cpu(), xr.grad) < 1e-3, report("dframes", x.grad.cpu(), xr.grad)
    assert rel_err(conv.weight.grad.cpu(), wr.grad) < 1e-3, report("dW", conv.weight.grad.cpu(), wr.grad)
    assert rel_err(conv.bias.grad.cpu(), br.grad) < 1e-3, report("db", conv.bias.grad.cpu(), br.grad)


def test_frameconv_loud_errors(cuda_device):
    from plconv import functional as PF
    conv = torch.nn.Conv2d(1, 32, 3, stride=2, padding=1).to(cuda_device)
    with pytest.raises(RuntimeError, match="contiguous fp32"):
        PF.frameconv(torch.zeros(2, 1, 8, 8, device=cuda_device, dtype=torch.bfloat16), conv)
    with pytest.raises(RuntimeError):
        PF.frameconv(torch.zeros(2, 1, 8, 8), conv)                                        # CPU tensor
    assert not PF.frameconv_supported(torch.nn.Conv2d(8, 32, 3, stride=2, padding=1))      # too many channels
    assert not PF.frameconv_supported(torch.nn.Conv2d(1, 24, 3, stride=2, padding=1))      # Cout / 8 not a power of two
    assert not PF.frameconv_supported(torch.nn.Conv2d(1, 32, 5, stride=2, padding=2))      # 5x5


def test_convnd_loud_errors(cuda_device):
    from plconv import functional as PF
    conv = torch.nn.Conv2d(8, 8, 4, stride=2, padding=2).to(cuda_device)          # even kernel
    with pytest.raises((RuntimeError, ValueError)):
        cp = PF.ConvNdParams(conv)
        PF.convnd(torch.zeros(1, 1, 8, 8, 8, device=cuda_device, dtype=torch.bfloat16), cp)
    conv = torch.nn.Conv2d(8, 8, 3, stride=3, padding=1).to(cuda_device)          # stride 3
    with pytest.raises(RuntimeError, match="strides"):
        PF.convnd(torch.zeros(1, 1, 9, 9, 8, device=cuda_device, dtype=torch.bfloat16), PF.ConvNdParams(conv))
    conv = torch.nn.Conv2d(8, 8, 3, stride=2, padding=1).to(cuda_device)
    with pytest.raises(RuntimeError, match="contiguous bf16"):
        PF.convnd(torch.zeros(1, 1, 8, 8, 8, device=cuda_device), PF.ConvNdParams(conv))   # fp32 input
    with pytest.raises(RuntimeError, match="no CPU path"):
        import plconv
        plconv.Discriminator()(torch.zeros(1, 4, 1, 16, 16))


def _spec_params(disc):
    return {k: v.detach().double().cpu() for k, v in disc.state_dict().items()}


def test_discriminator_matches_eager_spec_forward_and_gradients(cuda_device):
    """All four layers + logits vs oracle/gan_oracle.py (fp64 on the bf16-rounded clip), then d logits / d clip and
    every parameter gradient vs the spec's autograd."""
    import plconv
    from oracle import gan_oracle as G
    torch.manual_seed(3)
    disc = plconv.Discriminator().to(cuda_device)
    n, t, hh, ww = 3, 8, 32, 40
    clip = torch.relu(torch.randn(n, t, 1, hh, ww) + 0.3).to(torch.bfloat16).float()
    cd = clip.to(cuda_device).requires_grad_()
    a1, a2, a3, s = disc.features(cd)
    logits = s[..., 0].float().reshape(n, -1).mean(1)
    p = {k: v.requires_grad_() for k, v in _spec_params(disc).items()}
    cr = clip.double().requires_grad_()
    r1, r2, r3, rs = G.discriminator_features(cr, p)
    ref_logits = rs.reshape(n, -1).mean(1)
    assert rel_err(a1[:, 0].permute(0, 3, 1, 2).cpu(), r1) < 1e-2
    assert rel_err(a2.permute(0, 4, 1, 2, 3).cpu(), r2) < 1e-2
    assert rel_err(a3.permute(0, 4, 1, 2, 3).cpu(), r3) < 1.5e-2
    assert rel_err(s[:, 0, :, :, :1].permute(0, 3, 1, 2).cpu(), rs) < 2e-2
    assert rel_err(logits.cpu(), ref_logits) < 2e-2, (logits, ref_logits)
    # Gradients: the kernel takes every LeakyReLU mask from the bf16 activation it stored.  Where an activation is within
    # rounding error of 0 the fp64 spec may sit on the other side of the kink (gradient off by the slope factor on that
    # element: ~1e-3 of the elements per layer), so the spec's backward is conditioned on the SAME positive sets.
    masks = ((a1[:, 0] > 0).permute(0, 3, 1, 2).cpu(), (a2 > 0).permute(0, 4, 1, 2, 3).cpu(),
             (a3 > 0).permute(0, 4, 1, 2, 3).cpu())
    ref_logits = G.discriminator_features(cr, p, masks)[3].reshape(n, -1).mean(1)
    w = torch.tensor([1.0, -2.0, 0.5])
    (logits * w.to(cuda_device)).sum().backward()
    (ref_logits * w.double()).sum().backward()
    assert rel_err(cd.grad.cpu(), cr.grad) < 3e-2, report("dclip", cd.grad.cpu(), cr.grad)
    for k, v in disc.named_parameters():
        assert rel_err(v.grad.cpu(), p[k].grad) < 3e-2, report(k, v.grad.cpu(), p[k].grad)


def test_gan_losses_and_one_step_vs_eager_spec(cuda_device):
    """L_D, L1 and the adversarial term of a full rollout (generator + discriminator) vs the eager spec on identical
    weights (tolerance 2e-2 relative: bf16 activations through a 2 x 4-step rollout and four conv layers), then two
    optimisation steps: every parameter of G and D moves, stays finite, and the step never synchronises with the host."""
    import plconv
    from oracle import convlstm_oracle as O
    from oracle import gan_oracle as G
    torch.manual_seed(5)
    b, t_in, t_out, hh, ww, hd = 2, 4, 4, 32, 32, [16, 32]
    gen = plconv.NowcastGenerator(1, hd, 3, t_in, t_out, "bf16").to(cuda_device)
    disc = plconv.Discriminator().to(cuda_device)
    step = plconv.GanTrainStep(gen, disc, lambda_adv=0.05)
    frames = torch.relu(torch.randn(b, t_in, 1, hh, ww) + 0.3)
    target = torch.relu(torch.randn(b, t_out, 1, hh, ww) + 0.3)
    fd, td = frames.to(cuda_device), target.to(cuda_device)
    with torch.no_grad():
        fake = gen(fd)
        d_loss, l1, adv = step.losses(fd, td, fake)
    sd = {k: v.detach().double().cpu() for k, v in gen.state_dict().items()}
    L = len(hd)
    ref_fake = O.nowcast_forward(frames.double(), sd["init_conv.weight"], sd["init_conv.bias"],
                                 [sd[f"encoder.cells.{l}.conv.weight"] for l in range(L)],
                                 [sd[f"encoder.cells.{l}.conv.bias"] for l in range(L)],
                                 [sd[f"forecaster.cells.{l}.conv.weight"] for l in range(L)],
                                 [sd[f"forecaster.cells.{l}.conv.bias"] for l in range(L)],
                                 sd["head.weight"], sd["head.bias"], t_out)
    p = _spec_params(disc)
    f64, t64 = frames.double(), target.double()
    clips = torch.cat([torch.cat([f64, t64], 1), torch.cat([f64, ref_fake], 1)], 0)
    ref_d = G.d_loss(G.discriminator_forward(clips, p), b)
    ref_adv = G.g_adv_loss(G.discriminator_forward(torch.cat([f64, ref_fake], 1), p))
    ref_l1 = (ref_fake - t64).abs().mean()
    print(f"GAN losses  cuda: D {float(d_loss):.5f} L1 {float(l1):.5f} adv {float(adv):.5f}   "
          f"spec: D {float(ref_d):.5f} L1 {float(ref_l1):.5f} adv {float(ref_adv):.5f}")
    assert abs(float(d_loss) - float(ref_d)) < 2e-2 * abs(float(ref_d))
    assert abs(float(l1) - float(ref_l1)) < 2e-2 * abs(float(ref_l1))
    assert abs(float(adv) - float(ref_adv)) < 2e-2 * abs(float(ref_adv))

    before = [q.detach().clone() for q in list(gen.parameters()) + list(disc.parameters())]
    step(fd, td)                                       # warm-up (allocations, lazy init) outside the sync check
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        out = step(fd, td)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert torch.isfinite(out)
    for q0, q in zip(before, list(gen.parameters()) + list(disc.parameters())):
        assert torch.isfinite(q).all() and not torch.equal(q0, q)
    assert float(step.g.skipped_dev) == 0.0 and float(step.d.skipped_dev) == 0.0


def test_graphed_gan_step_matches_eager(cuda_device):
    """plconv.training.GraphedStep: the whole GAN step recorded as ONE CUDA graph (2 eager warm-up steps + capture +
    3 replays) vs 5 eager steps from the same seeds on the same batch: every step's generator loss within 2e-3 relative
    and every parameter of G and D within 2e-3 absolute (Adam steps of 5e-4 / 2e-4; the only non-determinism is the
    order of the fp32 `red.global.add` reductions in the weight-gradient kernels).  Also: an eager forward after the
    replays must see the CURRENT weights (packed-weight caches are invalidated by the replay)."""
    import plconv
    from plconv.training import GraphedStep
    b, t_in, t_out, hh, ww, hd = 2, 3, 3, 32, 32, [16, 16]
    torch.manual_seed(9)
    frames = torch.relu(torch.randn(b, t_in, 1, hh, ww, device=cuda_device) + 0.3)
    target = torch.relu(torch.randn(b, t_out, 1, hh, ww, device=cuda_device) + 0.3)

    def make():
        torch.manual_seed(21)
        gen = plconv.NowcastGenerator(1, hd, 3, t_in, t_out, "bf16").to(cuda_device)
        disc = plconv.Discriminator().to(cuda_device)
        return gen, disc, plconv.GanTrainStep(gen, disc, lambda_adv=0.05)

    gen_e, disc_e, step_e = make()
    eager_losses = [float(step_e(frames, target)) for _ in range(5)]
    gen_g, disc_g, step_g = make()
    graphed = GraphedStep(step_g, (frames, target), warmup=2)
    assert graphed.eager_steps == 2
    graph_losses = [float(graphed(frames, target).clone()) for _ in range(3)]
    print("eager losses", eager_losses, "graph replays", graph_losses)
    for a, g in zip(eager_losses[2:], graph_losses):
        assert abs(a - g) < 2e-3 * abs(a), (eager_losses, graph_losses)
    for (k, pe), pg in zip(list(gen_e.named_parameters()) + list(disc_e.named_parameters()),
                           list(gen_g.parameters()) + list(disc_g.parameters())):
        assert torch.isfinite(pg).all()
        assert (pe - pg).abs().max() < 2e-3, (k, float((pe - pg).abs().max()))
    with torch.no_grad():                              # stale packed weights would reproduce the pre-step prediction
        ye, yg = gen_e(frames), gen_g(frames)
    assert rel_err(yg.cpu(), ye.cpu()) < 2e-2, report("post-replay forward", yg.cpu(), ye.cpu())
    assert float(step_g.g.skipped_dev) == 0.0
