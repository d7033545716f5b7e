"""Parity at BASELINE.json's FULL sizes (configs[1] cell: B32 128x128 hidden 64; configs[3] cell: B16 256x256 hidden
128), where the CPU oracle would take minutes: size-independent properties plus windowed oracle checks.

 * locality: the oracle on a cropped window (+ conv halo) of a few samples must reproduce the interior of the
   full-size CUDA result (the cell is a local operator: convlstm.py:17-28)
 * two independent implementations: bf16 tensor-core kernel vs the fp32 SIMT kernel on the same inputs
 * batch-permutation equivariance, BIT-exact: tiles are decoded from a linear pixel index, so a sample's result must
   not depend on where in the batch it sits
 * backward linearity, bit-exact for the state gradients: (dh, dc) -> 2*(dh, dc) doubles dx / dh_prev / dc_prev
   exactly (power-of-two scaling commutes with every rounding); dW / db double within fp32 summation noise
 * zero-gradient: dh = dc = 0 gives exactly zero everywhere
"""
import pytest
import torch

from oracle import convlstm_oracle as O

pytestmark = pytest.mark.gpu

FULL = [("cfg2_cell", 32, 128, 128, 64, 64, 3), ("cfg4_radar_cell", 16, 256, 256, 128, 128, 3)]


def _mk(B, H, W, Cin, Ch, k, dev, seed=0):
    import plconv
    from plconv import functional as F
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = torch.randn(4 * Ch, Cin + Ch, k, k, generator=g) * (1.5 / ((Cin + Ch) * k * k) ** 0.5)
    b = torch.randn(4 * Ch, generator=g) * 0.2
    gd = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, H, W, Cin, device=dev, generator=gd).to(torch.bfloat16)
    h = torch.tanh(torch.randn(B, H, W, Ch, device=dev, generator=gd)).to(torch.bfloat16)
    c = torch.randn(B, H, W, Ch, device=dev, generator=gd)
    pw = F.pack_weights(w.to(dev), b.to(dev), Cin, Ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
    return plconv, F, w, b, x, h, c, pw


@pytest.mark.parametrize("name,B,H,W,Cin,Ch,k", FULL, ids=[f[0] for f in FULL])
def test_full_size_forward_windows_vs_oracle_and_fp32_kernel(name, B, H, W, Cin, Ch, k, cuda_device):
    plconv, F, w, b, x, h, c, pw = _mk(B, H, W, Cin, Ch, k, cuda_device)
    h2, c2 = F.cell_forward(x, h, c, pw)
    # (a) windows: corners (zero padding), an edge, the interior, first / last sample
    p = k // 2
    S = 24
    for (bi, y0, x0) in [(0, 0, 0), (B - 1, H - S, W - S), (B // 2, 0, W // 2 - S // 2), (1, H // 2, W // 3)]:
        ya, yb, xa, xb = max(0, y0 - p), min(H, y0 + S + p), max(0, x0 - p), min(W, x0 + S + p)
        crop = lambda t: t[bi:bi + 1, ya:yb, xa:xb].float().cpu().permute(0, 3, 1, 2).contiguous()
        ho, co = O.cell_forward(crop(x), crop(h), crop(c), w.to(torch.bfloat16).float(), b)
        # rows/cols of the crop whose receptive field lies inside the crop or at a true image border
        iy0 = 0 if ya == 0 else p
        iy1 = (yb - ya) if yb == H else (yb - ya) - p
        ix0 = 0 if xa == 0 else p
        ix1 = (xb - xa) if xb == W else (xb - xa) - p
        got_h = h2[bi, ya + iy0:ya + iy1, xa + ix0:xa + ix1].float().cpu().permute(2, 0, 1)
        got_c = c2[bi, ya + iy0:ya + iy1, xa + ix0:xa + ix1].float().cpu().permute(2, 0, 1)
        eh = float((got_h - ho[0, :, iy0:iy1, ix0:ix1]).abs().max())
        ec = float((got_c - co[0, :, iy0:iy1, ix0:ix1]).abs().max() / co.abs().max())
        assert eh <= 1e-2 and ec <= 1e-4, (name, bi, y0, x0, eh, ec)      # h: bf16 output rounding; c: fp32
    # (b) the independent fp32 SIMT kernel on a batch slice of the same inputs
    nb = 2
    pw32 = F.pack_weights(w.to(torch.bfloat16).float().to(cuda_device), b.to(cuda_device), Cin, Ch, k,
                          plconv.PLC_MODE_FP32)
    h32, c32 = F.cell_forward(x[:nb].float().contiguous(), h[:nb].float().contiguous(), c[:nb].contiguous(), pw32)
    assert float((h2[:nb].float() - h32).abs().max()) <= 1e-2
    assert float((c2[:nb] - c32).abs().max() / c32.abs().max()) <= 1e-4


@pytest.mark.parametrize("name,B,H,W,Cin,Ch,k", FULL[:1], ids=[FULL[0][0]])
def test_full_size_batch_permutation_is_bit_exact(name, B, H, W, Cin, Ch, k, cuda_device):
    plconv, F, w, b, x, h, c, pw = _mk(B, H, W, Cin, Ch, k, cuda_device, seed=1)
    h2, c2 = F.cell_forward(x, h, c, pw)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(5)).to(cuda_device)
    hp, cp = F.cell_forward(x[perm].contiguous(), h[perm].contiguous(), c[perm].contiguous(), pw)
    assert torch.equal(hp, h2[perm]) and torch.equal(cp, c2[perm])


@pytest.mark.parametrize("name,B,H,W,Cin,Ch,k", FULL[:1], ids=[FULL[0][0]])
def test_full_size_backward_linearity_and_zero(name, B, H, W, Cin, Ch, k, cuda_device):
    plconv, F, w, b, x, h, c, pw = _mk(B, H, W, Cin, Ch, k, cuda_device, seed=2)
    dev = cuda_device
    gd = torch.Generator(device=dev).manual_seed(9)
    dh = (torch.randn(B, H, W, Ch, device=dev, generator=gd) * 0.1).to(torch.bfloat16)
    dc = torch.randn(B, H, W, Ch, device=dev, generator=gd) * 0.1

    def bwd(dh_, dc_):
        dW = torch.zeros(4 * Ch, Cin + Ch, k, k, device=dev)
        db = torch.zeros(4 * Ch, device=dev)
        dx, dhp, dcp = F.cell_backward(x, h, c, pw, dh_, None, dc_, dW, db)
        return dx.clone(), dhp.clone(), dcp.clone(), dW, db

    a = bwd(dh, dc)
    b2 = bwd((dh.float() * 2).to(torch.bfloat16), dc * 2)
    for i, nm in enumerate(("dx", "dh_prev", "dc_prev")):
        assert torch.equal(b2[i].float(), a[i].float() * 2), nm
    for i, nm in ((3, "dW"), (4, "db")):
        err = float((b2[i] - 2 * a[i]).abs().max() / (2 * a[i]).abs().max())
        assert err <= 1e-5, (nm, err)
    z = bwd(torch.zeros_like(dh), torch.zeros_like(dc))
    assert all(float(t.abs().max()) == 0.0 for t in z)
    # dh2 (recurrent gradient) is summed inside the kernel: (dh, dh2) == (dh + dh2, None) when the sum is exact
    half = (dh.float() * 0.5).to(torch.bfloat16)
    dW = torch.zeros(4 * Ch, Cin + Ch, k, k, device=dev)
    db = torch.zeros(4 * Ch, device=dev)
    dx, dhp, dcp = F.cell_backward(x, h, c, pw, half, half, dc, dW, db)
    assert torch.equal(dx, a[0]) and torch.equal(dhp, a[1]) and torch.equal(dcp, a[2])


@pytest.mark.parametrize("name,B,H,W,Cin,Ch,k", FULL, ids=[f[0] for f in FULL])
def test_full_size_backward_windows_and_subbatch_wgrad_vs_oracle(name, B, H, W, Cin, Ch, k, cuda_device):
    """plc_cell_bwd at BASELINE's full cell sizes against oracle.cell_backward (fp64 on the same bf16-rounded operands):
      * dx, dh_prev (bf16 outputs of the dgrad contraction over K = 4Ch*k^2) and dc_prev (fp32) on windows -- corners
        (zero padding), an edge, the interior, first / last sample; the oracle runs on the window + 2*pad halo (gates need
        pad, the transposed conv of dZ needs pad more) and only pixels whose dependence cone lies inside the crop or
        ends at a true image border are compared;
      * dW, db of a 2-sample sub-batch against the oracle (fp64), a reduction over 2*H*W pixels;
      * the FULL-size dW, db (524 288 / 1 048 576-pixel split-K reduction with red.global.add) against the sum of the
        B/2 sub-batch results of the same kernel -- a different split of the same sum, so only fp32 summation-order noise
        may separate them.
    Tolerances (global-max norm, max|err| / max|ref| over the compared tensor): dx, dh_prev 1e-2 (bf16 dZ and bf16
    output rounding), dc_prev 1e-3, dW / db vs oracle 1e-2, full vs sum of sub-batches 5e-4 (fp32 accumulation of up to
    1 048 576 terms in two different orders; measured 1.9e-4 at the cfg4 size)."""
    plconv, F, w, b, x, h, c, pw = _mk(B, H, W, Cin, Ch, k, cuda_device, seed=3)
    dev = cuda_device
    gd = torch.Generator(device=dev).manual_seed(11)
    dh = (torch.randn(B, H, W, Ch, device=dev, generator=gd) * 0.1).to(torch.bfloat16)
    dc = torch.randn(B, H, W, Ch, device=dev, generator=gd) * 0.1
    dW = torch.zeros(4 * Ch, Cin + Ch, k, k, device=dev)
    db = torch.zeros(4 * Ch, device=dev)
    dx, dhp, dcp = F.cell_backward(x, h, c, pw, dh, None, dc, dW, db)
    wr = w.to(torch.bfloat16).double()
    p, S = k // 2, 20
    m = 2 * p
    for (bi, y0, x0) in [(0, 0, 0), (B - 1, H - S, W - S), (B // 2, 0, W // 2 - S // 2), (1, H // 2, W // 3)]:
        ya, yb, xa, xb = max(0, y0 - m), min(H, y0 + S + m), max(0, x0 - m), min(W, x0 + S + m)
        crop = lambda t: t[bi:bi + 1, ya:yb, xa:xb].double().cpu().permute(0, 3, 1, 2).contiguous()
        ref = O.cell_backward(crop(x), crop(h), crop(c), wr, b.double(), crop(dh), crop(dc))
        iy0 = 0 if ya == 0 else m
        iy1 = (yb - ya) if yb == H else (yb - ya) - m
        ix0 = 0 if xa == 0 else m
        ix1 = (xb - xa) if xb == W else (xb - xa) - m
        for nm, got, tol in (("dx", dx, 1e-2), ("dh_prev", dhp, 1e-2), ("dc_prev", dcp, 1e-3)):
            g_ = got[bi, ya + iy0:ya + iy1, xa + ix0:xa + ix1].double().cpu().permute(2, 0, 1)
            r_ = ref[nm][0, :, iy0:iy1, ix0:ix1]
            err = float((g_ - r_).abs().max() / r_.abs().max())
            assert err <= tol, (name, nm, bi, y0, x0, err)
    # sub-batch wgrad vs oracle
    nb = 2
    sub = lambda t: t[:nb].contiguous()
    dW2 = torch.zeros_like(dW)
    db2 = torch.zeros_like(db)
    F.cell_backward(sub(x), sub(h), sub(c), pw, sub(dh), None, sub(dc), dW2, db2)
    cpu = lambda t: t[:nb].double().cpu().permute(0, 3, 1, 2).contiguous()
    ref = O.cell_backward(cpu(x), cpu(h), cpu(c), wr, b.double(), cpu(dh), cpu(dc))
    for nm, got in (("dW", dW2), ("db", db2)):
        err = float((got.double().cpu() - ref[nm]).abs().max() / ref[nm].abs().max())
        assert err <= 1e-2, (name, nm, "sub-batch vs oracle", err)
    # full-size reduction == sum of sub-batch reductions
    accW, accb = dW2.double(), db2.double()
    for i in range(nb, B, nb):
        sl = lambda t: t[i:i + nb].contiguous()
        dWi = torch.zeros_like(dW)
        dbi = torch.zeros_like(db)
        F.cell_backward(sl(x), sl(h), sl(c), pw, sl(dh), None, sl(dc), dWi, dbi)
        accW += dWi.double()
        accb += dbi.double()
    for nm, got, acc in (("dW", dW, accW), ("db", db, accb)):
        err = float((got.double() - acc).abs().max() / acc.abs().max())
        assert err <= 5e-4, (name, nm, "full vs sum of sub-batches", err)


@pytest.mark.parametrize("name,B,H,W,Cin,Ch,k", FULL, ids=[f[0] for f in FULL])
def test_full_size_saved_gates_bptt_vs_recompute(name, B, H, W, Cin, Ch, k, cuda_device):
    """Saved-gates BPTT at the full cfg2 / cfg4 cell sizes (8192 / 8192 tiles, both 64-channel slices at hidden 128):
    forward outputs bit-identical to the plain forward, every gradient within the bf16 rounding of the stored gates of the
    recompute path (which the other tests of this file hold to the oracle): fp32 tensors <= 6e-3, bf16 tensors <= 1e-2 of
    max, and the MEAN deviation an order below that (no systematic tile / slice mix-up hides under a max norm)."""
    plconv, F, w, b, x, h, c, pw = _mk(B, H, W, Cin, Ch, k, cuda_device, seed=4)
    dev = cuda_device
    n = F.saved_gates_bytes(B, H, W, pw)
    assert n == B * H * W * 4 * Ch * 2                       # 8 bytes per hidden element (all tiles full at these sizes)
    saved = torch.empty(n, dtype=torch.uint8, device=dev)
    h_a, c_a = F.cell_forward(x, h, c, pw)
    h_b, c_b = F.cell_forward(x, h, c, pw, saved=saved)
    assert torch.equal(h_a, h_b) and torch.equal(c_a, c_b)
    gd = torch.Generator(device=dev).manual_seed(11)
    dh = torch.randn(B, H, W, Ch, device=dev, generator=gd).to(torch.bfloat16)
    dc = torch.randn(B, H, W, Ch, device=dev, generator=gd)

    def bwd(sv):
        img = F.wgrad_accumulator(B, H, W, pw, dev)
        db = torch.zeros(4 * Ch, device=dev)
        dx, dhp, dcp = F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, saved=sv)
        dW = torch.zeros(4 * Ch, Cin + Ch, k, k, device=dev)
        F.wgrad_unpack(img, pw, dW)
        return {"dx": dx, "dh_prev": dhp, "dc_prev": dcp, "dW": dW, "db": db}

    rec, sav = bwd(None), bwd(saved)
    for key in rec:
        a, s_ = rec[key].float(), sav[key].float()
        scale = float(a.abs().max())
        worst, mean = float((a - s_).abs().max()) / scale, float((a - s_).abs().mean()) / scale
        bf = rec[key].dtype == torch.bfloat16
        tol, mean_tol = (1e-2, 2e-3) if bf else (6e-3, 1.5e-3)       # measured: dx 5.5e-3 / 5e-4, db 2.6e-3 / 5e-4
        assert worst < tol and mean < mean_tol, (name, key, worst, mean)
