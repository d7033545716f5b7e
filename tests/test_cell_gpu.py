"""GPU parity: the CUDA cell step (through the C ABI) vs the CPU oracle and the reference goldens.

Tolerances (north_star): bf16 tensor-core mode <= 1e-2 relative (per-step h/c), fp32 validation
mode <= 1e-5.  "relative" = max |err| / max |ref| over the tensor (bf16 h has 2^-9 rounding).
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden
from oracle import convlstm_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2
FP32_TOL = 1e-5


def _plconv():
    import plconv
    from plconv import functional as F
    return plconv, F


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-30))


def report(name, got, ref):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    err = (got - ref).abs()
    idx = np.unravel_index(int(err.argmax()), err.shape)
    return (f"{name}: rel={float(err.max() / (ref.abs().max() + 1e-30)):.3e} max_abs={float(err.max()):.3e} at {idx} "
            f"got={float(got[idx]):.5f} ref={float(ref[idx]):.5f} frac>1e-2={(err > 1e-2).double().mean():.4f} "
            f"nan={int(torch.isnan(got).sum())}")


def nhwc(t, dtype, dev, c_pad=None):
    """[B,C,H,W] cpu tensor -> [B,H,W,C'] device tensor"""
    t = t.permute(0, 2, 3, 1)
    if c_pad is not None and c_pad != t.shape[-1]:
        t = torch.nn.functional.pad(t, (0, c_pad - t.shape[-1]))
    return t.contiguous().to(device=dev, dtype=dtype)


def nchw(t):
    return t.detach().float().cpu().permute(0, 3, 1, 2).contiguous()


def pad8(c):
    return (c + 7) // 8 * 8


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def run_cell(mode, x, h, c, w, b, dev):
    """x,h,c NCHW fp32 cpu; w OIHW; returns (h2, c2) NCHW fp32 cpu via the CUDA path."""
    plconv, F = _plconv()
    B, cin, H, W = x.shape if x is not None else (h.shape[0], 0, h.shape[2], h.shape[3])
    ch = h.shape[1]
    k = w.shape[-1]
    if mode == plconv.PLC_MODE_BF16_TC:
        cp = pad8(cin)
        adt = torch.bfloat16
    else:
        cp = cin
        adt = torch.float32
    pw = F.pack_weights(w.to(dev), None if b is None else b.to(dev), cin, ch, k, mode, cin_pad=cp)
    xd = nhwc(x, adt, dev, cp) if cin else None
    hd = nhwc(h, adt, dev)
    cd = nhwc(c, torch.float32, dev)
    h2, c2 = F.cell_forward(xd, hd, cd, pw)
    torch.cuda.synchronize()
    return nchw(h2), nchw(c2)


@pytest.mark.parametrize("path", golden_files("cell_"), ids=os.path.basename)
def test_fp32_mode_forward_vs_reference_golden(path, cuda_device):
    plconv, _ = _plconv()
    g = load_golden(path)
    T = torch.from_numpy
    h2, c2 = run_cell(plconv.PLC_MODE_FP32, T(g["x"]), T(g["h"]), T(g["c"]), T(g["weight"]), T(g["bias"]), cuda_device)
    eh, ec = rel_err(h2, T(g["h_next"])), rel_err(c2, T(g["c_next"]))
    assert eh < FP32_TOL and ec < FP32_TOL, report("h", h2, T(g["h_next"])) + " | " + report("c", c2, T(g["c_next"]))


@pytest.mark.parametrize("path", golden_files("cell_"), ids=os.path.basename)
def test_bf16_mode_forward_vs_reference_golden(path, cuda_device):
    plconv, _ = _plconv()
    g = load_golden(path)
    T = torch.from_numpy
    h2, c2 = run_cell(plconv.PLC_MODE_BF16_TC, T(g["x"]), T(g["h"]), T(g["c"]), T(g["weight"]), T(g["bias"]),
                      cuda_device)
    eh, ec = rel_err(h2, T(g["h_next"])), rel_err(c2, T(g["c_next"]))
    assert eh < BF16_TOL and ec < BF16_TOL, report("h", h2, T(g["h_next"])) + " | " + report("c", c2, T(g["c_next"]))


# (B, Cin, Ch, H, W, k)
SHAPES = [
    (1, 64, 64, 8, 16, 1),      # pure GEMM: one tile, one tap
    (1, 64, 64, 8, 16, 3),      # one tile, 9 taps, halo entirely zero padding
    (2, 64, 64, 16, 32, 3),     # several tiles with real halos
    (1, 16, 16, 12, 15, 3),     # ragged W, channel chunks < 64 (TMA OOB fill in C)
    (2, 8, 32, 9, 7, 5),        # k=5, odd sizes
    (1, 128, 128, 16, 16, 3),   # 2 N tiles (Ch=128), 2 K chunks per source
    (1, 32, 48, 10, 20, 3),     # CH_TILE=48
    (3, 0, 32, 8, 8, 3),        # no input tensor (forecaster first layer)
    (1, 72, 80, 6, 130, 3),     # W > 128, chunk tails, CH_TILE=16
    (2, 64, 256, 8, 8, 3),      # Ch = 256: 4 N tiles
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "B%d_Cin%d_Ch%d_%dx%d_k%d" % s)
@pytest.mark.parametrize("mode_name", ["bf16", "fp32"])
def test_forward_vs_oracle(shape, mode_name, cuda_device):
    plconv, _ = _plconv()
    B, cin, ch, H, W, k = shape
    mode = plconv.PLC_MODE_BF16_TC if mode_name == "bf16" else plconv.PLC_MODE_FP32
    gen = torch.Generator().manual_seed(hash(shape) % (2 ** 31))
    fan_in = (cin + ch) * k * k
    w = (torch.rand(4 * ch, cin + ch, k, k, generator=gen) * 2 - 1) * (3.0 / fan_in) ** 0.5 * 2
    b = torch.randn(4 * ch, generator=gen) * 0.5
    x = torch.randn(B, cin, H, W, generator=gen) if cin else None
    h = torch.randn(B, ch, H, W, generator=gen) * 0.5
    c = torch.randn(B, ch, H, W, generator=gen)
    if mode_name == "bf16":
        # isolate kernel error from operand quantisation: the oracle sees the same bf16-rounded operands
        xr, hr, wr = (None if x is None else bf16r(x)), bf16r(h), bf16r(w)
        tol = BF16_TOL
    else:
        xr, hr, wr = x, h, w
        tol = FP32_TOL
    h_ref, c_ref = O.cell_forward(None if xr is None else xr.double(), hr.double(), c.double(), wr.double(), b.double())
    h2, c2 = run_cell(mode, x, h, c, w, b, cuda_device)
    eh, ec = rel_err(h2, h_ref), rel_err(c2, c_ref)
    assert eh < tol and ec < tol, report("h", h2, h_ref) + " | " + report("c", c2, c_ref)
    # rel_err is a GLOBAL-max norm (max |err| / max |ref| over the tensor).  The cell state stays fp32 in both modes, so
    # it can also be held to a PER-ELEMENT bound: |err| <= atol + rtol * |ref| for every element -- bf16 mode: the
    # tanh.approx / ex2-based gates carry ~2^-11 relative error each (atol = rtol = 2e-3); fp32 mode: 1e-5.
    a = 2e-3 if mode_name == "bf16" else 1e-5
    err = (c2.double() - c_ref).abs()
    bound = a + a * c_ref.abs()
    worst = float((err / bound).max())
    assert worst <= 1.0, f"per-element c check: worst err/bound {worst:.2f}; " + report("c", c2, c_ref)


def test_kat_zero_weights_gpu(cuda_device):
    """W=0,b=0 => c'=0.5c, h'=0.5 tanh(0.5c); c=2 -> c'=1, h'=0.380797 (SURVEY 8c KAT)."""
    plconv, _ = _plconv()
    ch = 16
    w = torch.zeros(4 * ch, 8 + ch, 3, 3)
    x = torch.randn(1, 8, 5, 5)
    h = torch.randn(1, ch, 5, 5)
    c = torch.full((1, ch, 5, 5), 2.0)
    for mode, tol in ((plconv.PLC_MODE_FP32, 1e-6), (plconv.PLC_MODE_BF16_TC, 4e-3)):
        h2, c2 = run_cell(mode, x, h, c, w, torch.zeros(4 * ch), cuda_device)
        assert (c2 - 1.0).abs().max() < tol
        assert (h2 - 0.380797088).abs().max() < tol


def test_errors_are_loud(cuda_device):
    plconv, F = _plconv()
    dev = cuda_device
    w = torch.zeros(4 * 16, 8 + 16, 4, 4, device=dev)
    with pytest.raises(RuntimeError, match="odd"):
        F.pack_weights(w, None, 8, 16, 4, plconv.PLC_MODE_BF16_TC)
    w = torch.zeros(4 * 12, 8 + 12, 3, 3, device=dev)
    with pytest.raises(RuntimeError, match="Ch % 16"):
        F.pack_weights(w, None, 8, 12, 3, plconv.PLC_MODE_BF16_TC)
    # CPU tensors are refused: there is no CPU path
    pw = F.pack_weights(torch.zeros(64, 24, 3, 3, device=dev), None, 8, 16, 3, plconv.PLC_MODE_FP32)
    with pytest.raises(RuntimeError, match="CUDA"):
        F.cell_forward(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 16), torch.zeros(1, 4, 4, 16), pw)


# ---------------------------------------------------------------------------------------- backward

def run_cell_bwd(mode, x, h, c, w, b, gh, gc, dev):
    plconv, F = _plconv()
    B, cin, H, W = x.shape
    ch = h.shape[1]
    k = w.shape[-1]
    bf = mode == plconv.PLC_MODE_BF16_TC
    cp = pad8(cin) if bf else cin
    adt = torch.bfloat16 if bf else torch.float32
    pw = F.pack_weights(w.to(dev), b.to(dev), cin, ch, k, mode, with_dgrad=True, cin_pad=cp)
    xd, hd, cd = nhwc(x, adt, dev, cp), nhwc(h, adt, dev), nhwc(c, torch.float32, dev)
    dh, dc = nhwc(gh, adt, dev), nhwc(gc, torch.float32, dev)
    dW = torch.zeros(4 * ch, cp + ch, k, k, device=dev)
    db = torch.zeros(4 * ch, device=dev)
    dx, dhp, dcp = F.cell_backward(xd, hd, cd, pw, dh, None, dc, dW, db)
    torch.cuda.synchronize()
    dWc = dW.cpu()
    dW_ref_layout = torch.cat([dWc[:, :cin], dWc[:, cp:]], dim=1)
    return nchw(dx)[:, :cin], nchw(dhp), nchw(dcp), dW_ref_layout, db.cpu()


@pytest.mark.parametrize("path", golden_files("cell_"), ids=os.path.basename)
@pytest.mark.parametrize("mode_name", ["fp32", "bf16"])
def test_backward_vs_reference_autograd_golden(path, mode_name, cuda_device):
    plconv, _ = _plconv()
    g = load_golden(path)
    T = torch.from_numpy
    mode = plconv.PLC_MODE_BF16_TC if mode_name == "bf16" else plconv.PLC_MODE_FP32
    # fp32 mode: 1e-5 on the elementwise-dominated outputs, 5e-5 on the long fp32 reductions (dW, db: sums
    # over B*H*W pixels in a different order than ATen's).  bf16 mode: 2e-2 (bf16 dZ operand).
    tol = {"fp32": dict(dx=2e-5, dh_prev=2e-5, dc_prev=1e-5, dW=5e-5, db=5e-5),
           "bf16": dict(dx=2e-2, dh_prev=2e-2, dc_prev=1e-2, dW=2e-2, db=2e-2)}[mode_name]
    dx, dhp, dcp, dW, db = run_cell_bwd(mode, T(g["x"]), T(g["h"]), T(g["c"]), T(g["weight"]), T(g["bias"]),
                                        T(g["gh"]), T(g["gc"]), cuda_device)
    msgs = []
    for name, got in (("dx", dx), ("dh_prev", dhp), ("dc_prev", dcp), ("dW", dW), ("db", db)):
        ref = T(g[name])
        if rel_err(got, ref) >= tol[name]:
            msgs.append(report(name, got, ref))
    assert not msgs, " | ".join(msgs)


# ---------------------------------------------------------------------------------------- CTA-pair (cta_group::2) path
# Small shapes default to cta_group::1; force the pair path so that parity covers it too: odd number of 128-pixel
# tiles (tail tile of a pair decodes to image index B and must be zero-filled / clipped), ragged edges, several N
# tiles, narrow slices (direct-store epilogue) and 64-channel slices (TMA-store epilogue), no-input layer.
PAIR_SHAPES = [
    (1, 64, 64, 8, 16, 3),       # ONE tile: the pair's second CTA works on a fully out-of-range tile
    (3, 64, 64, 8, 16, 3),       # odd tile count
    (2, 64, 64, 16, 32, 3),
    (1, 16, 16, 12, 15, 3),      # ragged + narrow slice (direct stores)
    (1, 128, 128, 16, 16, 3),    # 2 N tiles
    (3, 0, 64, 8, 8, 3),         # no input tensor
    (2, 8, 32, 9, 7, 5),         # k = 5
    (1, 72, 80, 6, 130, 3),      # W > 128
]


@pytest.fixture
def force_pair():
    import plconv
    lib = plconv._lib.load()
    lib.plc_debug_set_cta_group(2)
    yield
    lib.plc_debug_set_cta_group(0)


@pytest.mark.parametrize("shape", PAIR_SHAPES, ids=lambda s: "B%d_Cin%d_Ch%d_%dx%d_k%d" % s)
def test_pair_path_forward_and_backward_vs_oracle(shape, force_pair, cuda_device):
    plconv, _ = _plconv()
    B, cin, ch, H, W, k = shape
    gen = torch.Generator().manual_seed(7 + sum(shape))
    fan_in = (cin + ch) * k * k
    w = (torch.rand(4 * ch, cin + ch, k, k, generator=gen) * 2 - 1) * (3.0 / fan_in) ** 0.5 * 2
    b = torch.randn(4 * ch, generator=gen) * 0.5
    x = torch.randn(B, max(cin, 1), H, W, generator=gen)[:, :cin] if cin else None
    h = torch.randn(B, ch, H, W, generator=gen) * 0.5
    c = torch.randn(B, ch, H, W, generator=gen)
    gh = torch.randn(B, ch, H, W, generator=gen)
    gc = torch.randn(B, ch, H, W, generator=gen)
    xr = None if x is None else bf16r(x).double()
    h_ref, c_ref = O.cell_forward(xr, bf16r(h).double(), c.double(), bf16r(w).double(), b.double())
    if cin:
        h2, c2 = run_cell(plconv.PLC_MODE_BF16_TC, x, h, c, w, b, cuda_device)
    else:
        from plconv import functional as F
        pw = F.pack_weights(w.to(cuda_device), b.to(cuda_device), 0, ch, k, plconv.PLC_MODE_BF16_TC)
        o = F.cell_forward(None, nhwc(h, torch.bfloat16, cuda_device), nhwc(c, torch.float32, cuda_device), pw)
        torch.cuda.synchronize()
        h2, c2 = nchw(o[0]), nchw(o[1])
    assert rel_err(h2, h_ref) < BF16_TOL and rel_err(c2, c_ref) < BF16_TOL, \
        report("h", h2, h_ref) + " | " + report("c", c2, c_ref)
    if cin:
        ref = O.cell_backward(xr, bf16r(h).double(), c.double(), bf16r(w).double(), b.double(), bf16r(gh).double(),
                              gc.double())
        dx, dhp, dcp, dW, db = run_cell_bwd(plconv.PLC_MODE_BF16_TC, x, h, c, w, b, gh, gc, cuda_device)
        bad = [report(n_, got, ref[n_]) for n_, got in (("dx", dx), ("dh_prev", dhp), ("dc_prev", dcp), ("dW", dW),
                                                         ("db", db)) if rel_err(got, ref[n_]) >= 2e-2]
        assert not bad, " | ".join(bad)


# ---- haloed-patch pipeline (one activation patch per tile serves all k*k taps as shifted UMMA views) ------------
# Forced ON (plc_debug_set_patch(1)) so that ragged / tiny images also take it; crossed with both cta_group paths.
PATCH_SHAPES = [
    (2, 64, 64, 16, 8, 3),       # exactly one 16x8 tile per image
    (2, 64, 64, 20, 13, 3),      # ragged in both directions
    (1, 128, 64, 33, 17, 3),     # 3 units (x: 2 chunks, h: 1)
    (3, 0, 64, 8, 8, 3),         # no input tensor, image smaller than a tile
    (1, 64, 64, 19, 21, 5),      # k = 5 (patch 20 x 12)
    (1, 64, 128, 16, 16, 3),     # 2 N tiles re-read the same patches
    (2, 72, 80, 18, 9, 3),       # channel tails inside the last 64-channel chunk
    (2, 0, 32, 20, 13, 3),       # single narrow source: 32-channel patch boxes (SWIZZLE_64B), 2 taps per K stage
    (1, 0, 16, 17, 9, 5),        # 16-channel patch boxes (SWIZZLE_32B), 4 taps per K stage, k = 5
]


@pytest.fixture(params=[1, 2], ids=["cta1", "cta2"])
def force_patch(request):
    import plconv
    lib = plconv._lib.load()
    lib.plc_debug_set_patch(1)
    lib.plc_debug_set_cta_group(request.param)
    yield
    lib.plc_debug_set_patch(-1)
    lib.plc_debug_set_cta_group(0)


@pytest.mark.parametrize("shape", PATCH_SHAPES, ids=lambda s: "B%d_Cin%d_Ch%d_%dx%d_k%d" % s)
def test_patch_pipeline_forward_and_backward_vs_oracle(shape, force_patch, cuda_device):
    test_pair_path_forward_and_backward_vs_oracle.__wrapped__(shape, None, cuda_device) \
        if hasattr(test_pair_path_forward_and_backward_vs_oracle, "__wrapped__") else \
        test_pair_path_forward_and_backward_vs_oracle(shape, None, cuda_device)


def test_patch_pipeline_matches_default_pipeline_closely(cuda_device):
    """Same inputs through both pipelines: identical products, different fp32 accumulation order only."""
    plconv, F = _plconv()
    lib = plconv._lib.load()
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    B, H, W, C = 4, 48, 40, 64
    w = torch.randn(4 * C, 2 * C, 3, 3, generator=g) * 0.05
    pw = F.pack_weights(w.to(dev), None, C, C, 3, plconv.PLC_MODE_BF16_TC)
    x = torch.randn(B, H, W, C, generator=g).to(dev).to(torch.bfloat16)
    h = torch.randn(B, H, W, C, generator=g).to(dev).to(torch.bfloat16)
    c = torch.randn(B, H, W, C, generator=g).to(dev)
    outs = []
    for mode in (0, 1):
        lib.plc_debug_set_patch(mode)
        try:
            h2, c2 = F.cell_forward(x, h, c, pw)
            outs.append((h2.float().clone(), c2.clone()))
        finally:
            lib.plc_debug_set_patch(-1)
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 2e-5
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 2 ** -7      # at most one bf16 ulp near |h| <= 1


@pytest.mark.parametrize("patch", [0, 1], ids=["default-pipeline", "patch-pipeline"])
@pytest.mark.parametrize("shape", [(2, 64, 64, 20, 13, 3), (1, 128, 64, 16, 16, 3), (2, 64, 128, 9, 17, 5),
                                   (2, 16, 16, 12, 12, 1), (1, 32, 32, 10, 10, 3)],
                         ids=lambda s: "B%d_Cin%d_Ch%d_%dx%d_k%d" % s)
def test_zero_state_form_is_bit_identical_to_zero_tensors(shape, patch, cuda_device):
    """plc_cell_fwd(h_prev = c_prev = NULL) (first step of every sequence, generator.py:156-160) == the same call on
    zero tensors, bit for bit; shapes the form does not cover must say so through plc_cell_fwd_zero_state_ok."""
    plconv, F = _plconv()
    lib = plconv._lib.load()
    B, cin, ch, H, W, k = shape
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(shape))
    w = torch.randn(4 * ch, cin + ch, k, k, generator=g) * 0.1
    b = torch.randn(4 * ch, generator=g) * 0.3
    pw = F.pack_weights(w.to(dev), b.to(dev), cin, ch, k, plconv.PLC_MODE_BF16_TC)
    x = torch.randn(B, H, W, cin, generator=g).to(dev).to(torch.bfloat16)
    lib.plc_debug_set_patch(patch)
    try:
        ok = F.zero_state_supported(pw)
        # 64-channel boxes always qualify; 16/32-channel boxes only if the x taps fill whole 64-element K stages
        # (1 tap of 16 channels, 9 taps of 32 channels: they do not)
        assert ok == (cin % 64 == 0)
        if not ok:
            with pytest.raises(RuntimeError, match="zero-state"):
                F.cell_forward_zero_state(x, pw)
            return
        h0 = torch.zeros(B, H, W, ch, device=dev, dtype=torch.bfloat16)
        c0 = torch.zeros(B, H, W, ch, device=dev)
        h_ref, c_ref = F.cell_forward(x, h0, c0, pw)
        h_z, c_z = F.cell_forward_zero_state(x, pw)
    finally:
        lib.plc_debug_set_patch(-1)
    assert torch.equal(h_z, h_ref) and torch.equal(c_z, c_ref)


def test_patch_vs_default_pipeline_random_shapes(cuda_device):
    """Differential sweep: 40 seeded random shapes (ragged / tiny / multi-tile images, 1-3 units, both N-tile counts,
    k = 3 and 5, both cta_group paths) through the haloed-patch pipeline and the default pipeline.  Same products,
    different fp32 accumulation order: c within 3e-5, h within one bf16 ulp, gradients within 1e-2 of their max."""
    plconv, F = _plconv()
    lib = plconv._lib.load()
    dev = cuda_device
    rng = np.random.RandomState(1234)
    for case in range(40):
        B = int(rng.randint(1, 4))
        H, W = int(rng.randint(1, 41)), int(rng.randint(1, 41))
        cin = int(rng.choice([0, 64, 128]))
        ch = int(rng.choice([64, 128]))
        k = int(rng.choice([3, 5]))
        cta = int(rng.choice([1, 2]))
        g = torch.Generator().manual_seed(case)
        w = torch.randn(4 * ch, cin + ch, k, k, generator=g) * (1.0 / ((cin + ch) * k * k) ** 0.5)
        b = torch.randn(4 * ch, generator=g) * 0.2
        pw = F.pack_weights(w.to(dev), b.to(dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
        x = torch.randn(B, H, W, cin, generator=g).to(dev).to(torch.bfloat16) if cin else None
        h = torch.randn(B, H, W, ch, generator=g).to(dev).to(torch.bfloat16)
        c = torch.randn(B, H, W, ch, generator=g).to(dev)
        dh = (torch.randn(B, H, W, ch, generator=g) * 0.1).to(dev).to(torch.bfloat16)
        dc = (torch.randn(B, H, W, ch, generator=g) * 0.1).to(dev)
        res = []
        lib.plc_debug_set_cta_group(cta)
        try:
            for mode in (0, 1):
                lib.plc_debug_set_patch(mode)
                h2, c2 = F.cell_forward(x, h, c, pw)
                dW = torch.zeros(4 * ch, cin + ch, k, k, device=dev)
                db = torch.zeros(4 * ch, device=dev)
                dx, dhp, dcp = F.cell_backward(x, h, c, pw, dh, None, dc, dW, db, need_dx=cin > 0)
                res.append([t.float().clone() for t in (h2, c2, dhp, dcp, dW, db) + ((dx,) if cin else ())])
        finally:
            lib.plc_debug_set_patch(-1)
            lib.plc_debug_set_cta_group(0)
        tag = f"case {case}: B{B} {H}x{W} Cin{cin} Ch{ch} k{k} cta{cta}"
        a, p_ = res
        assert float((a[0] - p_[0]).abs().max()) <= 2 ** -7, tag + " h"
        assert float((a[1] - p_[1]).abs().max()) <= 3e-5 * max(1.0, float(a[1].abs().max())), tag + " c"
        for i, nm in list(enumerate(("h", "c", "dh_prev", "dc_prev", "dW", "db", "dx")))[2:len(a)]:
            err = float((a[i] - p_[i]).abs().max() / (a[i].abs().max() + 1e-20))
            assert err <= 1e-2, f"{tag} {nm} {err:.3e}"


# ---- saved-gates BPTT: the forward keeps the activated gates, the backward gate kernel runs without its mainloop -----
SAVED_SHAPES = [
    (2, 64, 64, 16, 8, 3),       # exactly one tile per image
    (2, 64, 64, 20, 13, 3),      # ragged in both directions (uninitialised saved rows must be clipped)
    (3, 0, 64, 24, 24, 3),       # no input tensor (forecaster first layer)
    (1, 64, 128, 33, 17, 3),     # two 64-channel slices (n tiles)
    (2, 128, 64, 19, 21, 5),     # k = 5, two x chunks
]


@pytest.mark.parametrize("cta", [1, 2], ids=["cta1", "cta2"])
@pytest.mark.parametrize("shape", SAVED_SHAPES, ids=lambda s: "B%d_Cin%d_Ch%d_%dx%d_k%d" % s)
def test_saved_gates_bptt_matches_recompute_and_oracle(shape, cta, cuda_device):
    """plc_cell_fwd_save + plc_cell_bwd_saved: the forward outputs are bit-identical to plc_cell_fwd; the gradients
    agree with the recompute path up to the bf16 rounding of the stored gates (fp32 outputs <= 6e-3 of max, bf16 outputs
    <= 1e-2 = two ulps at the largest value) and with the fp64 oracle
    inside the same 2e-2 bound the recompute path is held to."""
    plconv, F = _plconv()
    lib = plconv._lib.load()
    lib.plc_debug_set_cta_group(cta)
    try:
        B, cin, ch, H, W, k = shape
        dev = cuda_device
        gen = torch.Generator().manual_seed(31 + sum(shape))
        fan_in = (cin + ch) * k * k
        w = (torch.rand(4 * ch, cin + ch, k, k, generator=gen) * 2 - 1) * (3.0 / fan_in) ** 0.5 * 2
        b = torch.randn(4 * ch, generator=gen) * 0.5
        x = torch.randn(B, cin, H, W, generator=gen) if cin else None
        h = torch.randn(B, ch, H, W, generator=gen) * 0.5
        c = torch.randn(B, ch, H, W, generator=gen)
        gh = torch.randn(B, ch, H, W, generator=gen)
        gh2 = torch.randn(B, ch, H, W, generator=gen) * 0.3
        gc = torch.randn(B, ch, H, W, generator=gen)
        pw = F.pack_weights(w.to(dev), b.to(dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
        nbytes = F.saved_gates_bytes(B, H, W, pw)
        assert nbytes > 0 and nbytes % (64 * 1024) == 0
        xd = nhwc(x, torch.bfloat16, dev) if cin else None
        hd, cd = nhwc(h, torch.bfloat16, dev), nhwc(c, torch.float32, dev)
        saved = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device=dev)        # bf16 NaN pattern where nothing is stored
        h_a, c_a = F.cell_forward(xd, hd, cd, pw)
        h_b, c_b = F.cell_forward(xd, hd, cd, pw, saved=saved)
        assert torch.equal(h_a, h_b) and torch.equal(c_a, c_b)
        dh, dh2, dc = nhwc(gh, torch.bfloat16, dev), nhwc(gh2, torch.bfloat16, dev), nhwc(gc, torch.float32, dev)

        def bwd(sv):
            img = F.wgrad_accumulator(B, H, W, pw, dev)
            db = torch.zeros(4 * ch, device=dev)
            dx, dhp, dcp = F.cell_backward_acc(xd, hd, cd, pw, dh, dh2, dc, img, db, saved=sv)
            dW = torch.zeros(4 * ch, cin + ch, k, k, device=dev)
            F.wgrad_unpack(img, pw, dW)
            return dx, dhp, dcp, dW, db

        rec, sav = bwd(None), bwd(saved)
        torch.cuda.synchronize()
        names = ("dx", "dh_prev", "dc_prev", "dW", "db")
        for n_, a, s_ in zip(names, rec, sav):
            if a is None:
                assert s_ is None
                continue
            assert torch.isfinite(s_.float()).all(), n_
            # dx / dh_prev are bf16 tensors: one or two ulps at the largest value are 4e-3 .. 8e-3 of max
            tol = 1e-2 if s_.dtype == torch.bfloat16 else 6e-3
            assert rel_err(s_, a) < tol, report(n_ + " saved vs recompute", s_, a)
        xr = None if x is None else bf16r(x).double()
        ref = O.cell_backward(xr, bf16r(h).double(), c.double(), bf16r(w).double(), b.double(),
                              (bf16r(gh) + bf16r(gh2)).double(), gc.double())
        got = {"dx": None if sav[0] is None else nchw(sav[0]), "dh_prev": nchw(sav[1]), "dc_prev": nchw(sav[2]),
               "dW": sav[3].cpu(), "db": sav[4].cpu()}
        bad = [report(n_, got[n_], ref[n_]) for n_ in names if got[n_] is not None and rel_err(got[n_], ref[n_]) >= 2e-2]
        assert not bad, " | ".join(bad)
    finally:
        lib.plc_debug_set_cta_group(0)


def test_saved_gates_unsupported_shapes_are_loud(cuda_device):
    plconv, F = _plconv()
    dev = cuda_device
    w = torch.randn(4 * 32, 32 + 32, 3, 3, device=dev) * 0.05
    pw = F.pack_weights(w, None, 32, 32, 3, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
    assert F.saved_gates_bytes(2, 16, 16, pw) == 0                              # Ch % 64 != 0: recompute only
    x = torch.zeros(2, 16, 16, 32, device=dev, dtype=torch.bfloat16)
    c = torch.zeros(2, 16, 16, 32, device=dev)
    with pytest.raises(RuntimeError, match="saved-gates"):
        F.cell_forward(x, x.clone(), c, pw, saved=torch.empty(65536, dtype=torch.uint8, device=dev))
