"""Tensor-level wrappers over the C ABI (torch is plumbing: device memory + current stream).

Working layout: NHWC-contiguous tensors of shape [B, H, W, C].
  bf16 tensor-core mode : x, h bf16; c fp32.     fp32 validation mode: everything fp32.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import PLC_MODE_BF16_TC, PLC_MODE_FP32, PLC_PACK_DGRAD, PLC_PACK_FWD, PlcCellDesc

Tensor = torch.Tensor


def _ptr(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev=None):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _call(t: Tensor, fn, what: str, *args) -> None:
    """Run one launching ABI call with `t`'s device current and that device's current stream appended as the last
    argument.  The library launches on the CURRENT device, so a tensor living elsewhere (``Trainer(device='cuda:1')``
    without ``torch.cuda.set_device``) must switch devices around the call -- never launch on cuda:0 with cuda:1
    pointers."""
    dev = t.device
    if dev.index == torch.cuda.current_device():
        _lib.check(fn(*args, _stream(dev)), what)
    else:
        with torch.cuda.device(dev):
            _lib.check(fn(*args, _stream(dev)), what)


def _require_cuda(t: Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this library has no CPU path")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous NHWC [B,H,W,C]")


def make_desc(B: int, H: int, W: int, Cin: int, Ch: int, k: int, mode: int, has_bias: bool) -> PlcCellDesc:
    return PlcCellDesc(B, H, W, Cin, Ch, k, mode, 1 if has_bias else 0)


@dataclass
class PackedWeights:
    """Kernel-ready images of one cell's `conv.weight` (+ the fp32 bias in reference order)."""
    fwd: Tensor
    dgrad: Optional[Tensor]
    bias: Optional[Tensor]
    mode: int
    Cin: int
    Ch: int
    k: int


def pack_weights(weight: Tensor, bias: Optional[Tensor], Cin: int, Ch: int, k: int, mode: int,
                 with_dgrad: bool = False, cin_pad: Optional[int] = None) -> PackedWeights:
    """weight: reference `conv.weight` [4Ch, Cin+Ch, k, k] (any float dtype, CUDA).

    cin_pad: pad the x-channel block with zero columns up to this many channels (bf16 mode needs
    Cin % 8 == 0; the activations are padded the same way by the caller)."""
    lib = _lib.load()
    assert weight.shape == (4 * Ch, Cin + Ch, k, k), (tuple(weight.shape), Cin, Ch, k)
    w = weight.detach().to(torch.float32)
    if cin_pad is not None and cin_pad != Cin:
        wp = torch.zeros(4 * Ch, cin_pad + Ch, k, k, device=w.device, dtype=torch.float32)
        wp[:, :Cin] = w[:, :Cin]
        wp[:, cin_pad:] = w[:, Cin:]
        w, Cin = wp, cin_pad
    w = w.contiguous()
    d = make_desc(1, 1, 1, Cin, Ch, k, mode, bias is not None)
    out = {}
    for kind, want in ((PLC_PACK_FWD, True), (PLC_PACK_DGRAD, with_dgrad)):
        if not want:
            out[kind] = None
            continue
        nbytes = lib.plc_packed_weight_bytes(ctypes.byref(d), kind)
        if nbytes == 0:
            _lib.check(-1, "plc_packed_weight_bytes")
        buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        _call(w, lib.plc_pack_weight, "plc_pack_weight", ctypes.byref(d), kind, _ptr(w), _ptr(buf))
        out[kind] = buf
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedWeights(out[PLC_PACK_FWD], out[PLC_PACK_DGRAD], b, mode, Cin, Ch, k)


def _act_dtype(mode: int):
    return torch.bfloat16 if mode == PLC_MODE_BF16_TC else torch.float32


def saved_gates_bytes(B: int, H: int, W: int, pw: PackedWeights) -> int:
    """Bytes of one step's saved-gates buffer (plc_saved_gates_bytes); 0 = this shape/mode has no saved-gates form."""
    lib = _lib.load()
    d = make_desc(B, H, W, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    return int(lib.plc_saved_gates_bytes(ctypes.byref(d)))


def cell_forward(x: Optional[Tensor], h: Tensor, c: Tensor, pw: PackedWeights,
                 h_out: Optional[Tensor] = None, c_out: Optional[Tensor] = None,
                 gates_out: Optional[Tensor] = None, saved: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """One fused cell step (convlstm.py:16-28).  NHWC tensors; returns (h_next, c_next).  `saved` (uint8,
    :func:`saved_gates_bytes`) additionally keeps the activated gates for :func:`cell_backward_acc`'s saved form."""
    lib = _lib.load()
    B, H, W, Ch = h.shape
    adt = _act_dtype(pw.mode)
    _require_cuda(h, "h"); _require_cuda(c, "c")
    if h.dtype != adt or c.dtype != torch.float32:
        raise RuntimeError(f"dtype mismatch: h {h.dtype} (want {adt}), c {c.dtype} (want float32)")
    if pw.Cin > 0:
        _require_cuda(x, "x")
        if x.dtype != adt or tuple(x.shape) != (B, H, W, pw.Cin):
            raise RuntimeError(f"x must be {adt} [B,H,W,{pw.Cin}], got {x.dtype} {tuple(x.shape)}")
    if Ch != pw.Ch or tuple(c.shape) != (B, H, W, Ch):
        raise RuntimeError("state shape mismatch")
    if h_out is None:
        h_out = torch.empty_like(h)
    if c_out is None:
        c_out = torch.empty_like(c)
    d = make_desc(B, H, W, pw.Cin, Ch, pw.k, pw.mode, pw.bias is not None)
    if saved is not None:
        if gates_out is not None:
            raise RuntimeError("cell_forward: pass gates_out or saved, not both")
        _call(h, lib.plc_cell_fwd_save, "plc_cell_fwd_save", ctypes.byref(d), _ptr(x) if pw.Cin > 0 else None, _ptr(h), _ptr(c),
              _ptr(pw.fwd), _ptr(pw.bias), _ptr(h_out), _ptr(c_out), _ptr(saved))
        return h_out, c_out
    _call(h, lib.plc_cell_fwd, "plc_cell_fwd", ctypes.byref(d), _ptr(x) if pw.Cin > 0 else None, _ptr(h), _ptr(c), _ptr(pw.fwd),
                                _ptr(pw.bias), _ptr(h_out), _ptr(c_out), _ptr(gates_out))
    return h_out, c_out


def zero_state_supported(pw: PackedWeights) -> bool:
    """True if :func:`cell_forward_zero_state` is available for these weights (plc_cell_fwd_zero_state_ok)."""
    lib = _lib.load()
    d = make_desc(1, 16, 16, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    return bool(lib.plc_cell_fwd_zero_state_ok(ctypes.byref(d)))


def cell_forward_zero_state(x: Tensor, pw: PackedWeights, h_out: Optional[Tensor] = None,
                            c_out: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """First step of a sequence: h_prev = c_prev = 0 (generator.py:156-160) without materialising the zero tensors --
    the h taps (half of the K loop) and the c_prev read are skipped.  Bit-identical to :func:`cell_forward` on zeros."""
    lib = _lib.load()
    _require_cuda(x, "x")
    B, H, W, cin = x.shape
    adt = _act_dtype(pw.mode)
    if x.dtype != adt or cin != pw.Cin:
        raise RuntimeError(f"x must be {adt} [B,H,W,{pw.Cin}], got {x.dtype} {tuple(x.shape)}")
    if h_out is None:
        h_out = torch.empty(B, H, W, pw.Ch, dtype=adt, device=x.device)
    if c_out is None:
        c_out = torch.empty(B, H, W, pw.Ch, dtype=torch.float32, device=x.device)
    d = make_desc(B, H, W, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    _call(x, lib.plc_cell_fwd, "plc_cell_fwd (zero state)", ctypes.byref(d), _ptr(x), None, None, _ptr(pw.fwd), _ptr(pw.bias), _ptr(h_out),
                                _ptr(c_out), None)
    return h_out, c_out


def bwd_workspace_bytes(B: int, H: int, W: int, pw: PackedWeights) -> int:
    lib = _lib.load()
    d = make_desc(B, H, W, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    return int(lib.plc_bwd_workspace_bytes(ctypes.byref(d)))


def bwd_workspace(B: int, H: int, W: int, pw: PackedWeights, device) -> Tensor:
    lib = _lib.load()
    d = make_desc(B, H, W, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    return torch.empty(lib.plc_bwd_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device=device)


def wgrad_accumulator(B: int, H: int, W: int, pw: PackedWeights, device) -> Tensor:
    """Zeroed fp32 accumulator image for dW (layout private to the library; see plc_wgrad_acc_bytes)."""
    lib = _lib.load()
    d = make_desc(B, H, W, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    return torch.zeros(lib.plc_wgrad_acc_bytes(ctypes.byref(d)) // 4, dtype=torch.float32, device=device)


def wgrad_unpack(acc: Tensor, pw: PackedWeights, dW: Tensor) -> Tensor:
    """dW [4Ch, Cin+Ch, k, k] fp32 (reference layout, Cin = pw.Cin) += unpack(acc)."""
    lib = _lib.load()
    d = make_desc(1, 1, 1, pw.Cin, pw.Ch, pw.k, pw.mode, pw.bias is not None)
    _call(acc, lib.plc_wgrad_unpack, "plc_wgrad_unpack", ctypes.byref(d), _ptr(acc), _ptr(dW))
    return dW


def cell_backward_acc(x: Optional[Tensor], h_prev: Tensor, c_prev: Tensor, pw: PackedWeights, dh: Tensor,
                      dh2: Optional[Tensor], dc_next: Optional[Tensor], dW_img: Optional[Tensor],
                      db_acc: Optional[Tensor], need_dx: bool = True, workspace: Optional[Tensor] = None,
                      dx: Optional[Tensor] = None, dh_prev: Optional[Tensor] = None,
                      dc_prev: Optional[Tensor] = None, saved: Optional[Tensor] = None):
    """BPTT of one cell step: returns (dx, dh_prev, dc_prev).  `dW_img` is the accumulator image from
    :func:`wgrad_accumulator` (accumulated in place across steps; convert once with :func:`wgrad_unpack`);
    `db_acc` [4Ch] fp32 is accumulated in place.  `saved` = the buffer the forward step filled (``cell_forward(...,
    saved=)``): the gate recompute contraction is skipped (plc_cell_bwd_saved)."""
    lib = _lib.load()
    B, H, W, Ch = h_prev.shape
    if pw.dgrad is None:
        raise RuntimeError("weights were packed without the dgrad image (with_dgrad=True)")
    if workspace is None:
        workspace = bwd_workspace(B, H, W, pw, h_prev.device)
    if dx is None and need_dx and pw.Cin > 0:
        dx = torch.empty_like(x)
    if dh_prev is None:
        dh_prev = torch.empty_like(h_prev)
    if dc_prev is None:
        dc_prev = torch.empty_like(c_prev)
    d = make_desc(B, H, W, pw.Cin, Ch, pw.k, pw.mode, pw.bias is not None)
    if saved is not None:
        _call(h_prev, lib.plc_cell_bwd_saved, "plc_cell_bwd_saved", ctypes.byref(d), _ptr(x) if pw.Cin > 0 else None,
              _ptr(h_prev), _ptr(c_prev), _ptr(saved), _ptr(pw.dgrad), _ptr(dh), _ptr(dh2), _ptr(dc_next),
              _ptr(dx), _ptr(dh_prev), _ptr(dc_prev), _ptr(dW_img), _ptr(db_acc), _ptr(workspace), workspace.numel())
        return dx, dh_prev, dc_prev
    _call(h_prev, lib.plc_cell_bwd, "plc_cell_bwd", ctypes.byref(d), _ptr(x) if pw.Cin > 0 else None, _ptr(h_prev), _ptr(c_prev),
                                _ptr(pw.fwd), _ptr(pw.dgrad), _ptr(pw.bias), _ptr(dh), _ptr(dh2), _ptr(dc_next),
                                _ptr(dx), _ptr(dh_prev), _ptr(dc_prev), _ptr(dW_img), _ptr(db_acc),
                                _ptr(workspace), workspace.numel())
    return dx, dh_prev, dc_prev


def cell_wgrad(x: Optional[Tensor], h_prev: Tensor, dz: Tensor, pw: PackedWeights, dW_img: Tensor,
               db_acc: Optional[Tensor]) -> None:
    """dW / db of MANY cell steps in one launch (plc_cell_wgrad): `h_prev` [N,H,W,Ch], `x` [N,H,W,Cin] and `dz` (the
    per-step workspaces of :func:`cell_backward_acc` called with ``dW_img=None``, N*H*W*4Ch elements) stacked over the
    steps, N = T*B.  Accumulates into `dW_img` / `db_acc` like the per-step form."""
    lib = _lib.load()
    N, H, W, Ch = h_prev.shape
    d = make_desc(N, H, W, pw.Cin, Ch, pw.k, pw.mode, pw.bias is not None)
    _call(h_prev, lib.plc_cell_wgrad, "plc_cell_wgrad", ctypes.byref(d), _ptr(x) if pw.Cin > 0 else None, _ptr(h_prev),
          _ptr(dz), _ptr(dW_img), _ptr(db_acc))


def cell_backward(x: Optional[Tensor], h_prev: Tensor, c_prev: Tensor, pw: PackedWeights, dh: Tensor,
                  dh2: Optional[Tensor], dc_next: Optional[Tensor], dW_acc: Optional[Tensor],
                  db_acc: Optional[Tensor], **kw):
    """Single-step convenience form: `dW_acc` is fp32 in the REFERENCE layout [4Ch, Cin+Ch, k, k] (Cin = pw.Cin, i.e.
    the padded count) and is += in place (accumulator image + unpack handled here)."""
    B, H, W, _ = h_prev.shape
    img = wgrad_accumulator(B, H, W, pw, h_prev.device) if dW_acc is not None else None
    out = cell_backward_acc(x, h_prev, c_prev, pw, dh, dh2, dc_next, img, db_acc, **kw)
    if dW_acc is not None:
        wgrad_unpack(img, pw, dW_acc)
    return out


def nchw_to_nhwc(src: Tensor, mode: int, c_pad: Optional[int] = None) -> Tensor:
    """Reference layout [B,C,H,W] fp32 -> working layout [B,H,W,C'] (bf16 in TC mode, zero-padded to c_pad)."""
    B, C, H, W = src.shape
    cd = C if c_pad is None else c_pad
    if mode == PLC_MODE_FP32:
        out = src.to(torch.float32).permute(0, 2, 3, 1)
        if cd != C:
            out = torch.nn.functional.pad(out, (0, cd - C))
        return out.contiguous()
    lib = _lib.load()
    s = src.to(torch.float32).contiguous()
    out = torch.empty(B, H, W, cd, dtype=torch.bfloat16, device=src.device)
    _call(s, lib.plc_nchw_f32_to_nhwc_bf16, "plc_nchw_f32_to_nhwc_bf16", _ptr(s), _ptr(out), B, C, cd, H, W)
    return out


def nhwc_to_nchw(src: Tensor) -> Tensor:
    """Working layout [B,H,W,C] -> reference layout [B,C,H,W] fp32."""
    B, H, W, C = src.shape
    if src.dtype == torch.float32:
        return src.permute(0, 3, 1, 2).contiguous()
    lib = _lib.load()
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=src.device)
    _call(src, lib.plc_nhwc_bf16_to_nchw_f32, "plc_nhwc_bf16_to_nchw_f32", _ptr(src.contiguous()), _ptr(out), B, C, H, W)
    return out


def frontend_forward(frames: Tensor, weight: Tensor, bias: Optional[Tensor], mode: int,
                     c_stride: Optional[int] = None, out: Optional[Tensor] = None) -> Tensor:
    """relu(init_conv(add_coord_channels(frames))) (generator.py:166-168): frames [N,Cf,H,W] fp32 ->
    features [N,H,W,c_stride] in the working layout (channels >= C untouched / zero)."""
    lib = _lib.load()
    N, Cf, H, W = frames.shape
    C = weight.shape[0]
    assert tuple(weight.shape) == (C, Cf + 2, 3, 3), tuple(weight.shape)
    cs = C if c_stride is None else c_stride
    _require_cuda(frames, "frames")
    if frames.dtype != torch.float32:
        raise RuntimeError("frames must be float32")
    if out is None:
        out = (torch.zeros if cs != C else torch.empty)(N, H, W, cs, dtype=_act_dtype(mode), device=frames.device)
    w = weight.detach().to(torch.float32).contiguous()
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    _call(frames, lib.plc_frontend_fwd, "plc_frontend_fwd", _ptr(frames), N, Cf, H, W, _ptr(w), _ptr(b), C, cs, mode, _ptr(out))
    return out


def head_forward(h: Tensor, weight: Tensor, bias: Optional[Tensor], mode: int, out: Optional[Tensor] = None) -> Tensor:
    """1x1 conv C -> 1 over working-layout h [..., C]; returns fp32 [...] (one value per pixel)."""
    lib = _lib.load()
    C = h.shape[-1]
    npix = h.numel() // C
    _require_cuda(h, "h")
    w = weight.detach().to(torch.float32).reshape(-1).contiguous()
    assert w.numel() == C
    b = None if bias is None else bias.detach().to(torch.float32).reshape(-1).contiguous()
    if out is None:
        out = torch.empty(h.shape[:-1], dtype=torch.float32, device=h.device)
    _call(h, lib.plc_head_fwd, "plc_head_fwd", _ptr(h), npix, C, _ptr(w), _ptr(b), mode, _ptr(out))
    return out


class _HeadFn(torch.autograd.Function):
    """Differentiable 1x1 head on working-layout h [..., C] -> fp32 [...] (plc_head_fwd / plc_head_bwd)."""

    @staticmethod
    def forward(ctx, h, weight, bias, mode):
        out = head_forward(h, weight, bias, mode)
        ctx.save_for_backward(h, weight)
        ctx.has_bias, ctx.mode = bias is not None, mode
        return out

    @staticmethod
    def backward(ctx, dy):
        h, weight = ctx.saved_tensors
        lib = _lib.load()
        C = h.shape[-1]
        npix = h.numel() // C
        if ctx.mode != PLC_MODE_BF16_TC:
            w = weight.reshape(-1).to(torch.float32)
            dyf = dy.to(torch.float32)
            dh = dyf.unsqueeze(-1) * w
            dw = (dyf.reshape(-1, 1) * h.reshape(-1, C).to(torch.float32)).sum(0)
            return dh.to(h.dtype), dw.reshape(weight.shape).to(weight.dtype), (dyf.sum().reshape(1) if ctx.has_bias else None), None
        dyc = dy.to(torch.float32).contiguous()
        w = weight.detach().to(torch.float32).reshape(-1).contiguous()
        dh = torch.empty_like(h)
        dw = torch.zeros(C, dtype=torch.float32, device=h.device)
        db = torch.zeros(1, dtype=torch.float32, device=h.device) if ctx.has_bias else None
        _call(h, lib.plc_head_bwd, "plc_head_bwd", _ptr(h), npix, C, _ptr(w), _ptr(dyc), _ptr(dh), _ptr(dw), _ptr(db))
        return dh, dw.reshape(weight.shape).to(weight.dtype), db, None


def head(h: Tensor, weight: Tensor, bias: Optional[Tensor], mode: int) -> Tensor:
    """Differentiable 1x1 head: h [..., C] (contiguous, working layout) -> fp32 [...]."""
    return _HeadFn.apply(h.contiguous(), weight, bias, mode)


# ------------------------------------------------------------------------------------------------ generic conv (bf16)
from ._lib import PlcConvDesc  # noqa: E402


def _rup(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class ConvParams:
    """Kernel-ready images of one plain conv layer's parameters, cached per parameter version.

    Wraps the reference's `nn.Conv2d` PARAMETER HOLDER (weight [Cout, Cin, k, k], bias [Cout]); channel counts are
    zero-padded to the kernel's granularity (Cin -> multiple of 8, Cout -> multiple of 8, or 32 with PixelShuffle)."""

    def __init__(self, conv: torch.nn.Conv2d, relu: bool = False, pixel_shuffle: bool = False, out_f32: bool = False):
        if out_f32 and (relu or pixel_shuffle):
            raise ValueError("out_f32 is for a final linear layer: no ReLU, no PixelShuffle")
        self.conv, self.relu, self.shuffle, self.out_f32 = conv, relu, pixel_shuffle, out_f32
        self.Cout, self.Cin, self.k, _ = conv.weight.shape
        self.cin_p = _rup(self.Cin, 8)
        self.cout_p = _rup(self.Cout, 32 if pixel_shuffle else 8)
        self._cache = None

    def desc(self, B, H, W) -> PlcConvDesc:
        return PlcConvDesc(B, H, W, self.cin_p, self.cout_p, self.k, int(self.relu), int(self.shuffle),
                           int(self.conv.bias is not None))

    def packed(self, need_dgrad: bool):
        w, b = self.conv.weight, self.conv.bias
        key = (_lib.weight_generation(), w.data_ptr(), w._version,
               None if b is None else (b.data_ptr(), b._version), str(w.device))
        pc = self._cache
        if pc is not None and pc[0] == key and (pc[3] is not None or not need_dgrad):
            return pc[1], pc[2], pc[3]
        lib = _lib.load()
        wp = torch.zeros(self.cout_p, self.cin_p, self.k, self.k, device=w.device, dtype=torch.float32)
        wp[:self.Cout, :self.Cin] = w.detach().to(torch.float32)
        bp = None
        if b is not None:
            bp = torch.zeros(self.cout_p, device=w.device, dtype=torch.float32)
            bp[:self.Cout] = b.detach().to(torch.float32)
        d = self.desc(1, 1, 1)
        fwd = torch.empty(lib.plc_conv_packed_weight_bytes(ctypes.byref(d), PLC_PACK_FWD), dtype=torch.uint8,
                          device=w.device)
        bias_packed = torch.zeros(self.cout_p, device=w.device, dtype=torch.float32)
        _call(wp, lib.plc_conv_pack_weight, "plc_conv_pack_weight", ctypes.byref(d), PLC_PACK_FWD, _ptr(wp), _ptr(bp), _ptr(fwd),
                                            _ptr(bias_packed))
        dg = None
        if need_dgrad:
            dg = torch.empty(lib.plc_conv_packed_weight_bytes(ctypes.byref(d), PLC_PACK_DGRAD), dtype=torch.uint8,
                             device=w.device)
            _call(wp, lib.plc_conv_pack_weight, "plc_conv_pack_weight", ctypes.byref(d), PLC_PACK_DGRAD, _ptr(wp), None, _ptr(dg), None)
        self._cache = (key, fwd, bias_packed, dg)
        return fwd, bias_packed, dg


class _ConvFn(torch.autograd.Function):
    """conv 'same' (+bias, +PixelShuffle(2), +ReLU) on NHWC bf16 tensors through plc_conv_fwd / plc_conv_bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias, cp: ConvParams):
        lib = _lib.load()
        B, H, W, C = x.shape
        if C != cp.cin_p or x.dtype != torch.bfloat16 or not x.is_contiguous():
            raise RuntimeError(f"conv input must be contiguous bf16 [B,H,W,{cp.cin_p}], got {x.dtype} {tuple(x.shape)}")
        # (torch.is_grad_enabled() is always False inside forward; ctx.needs_input_grad says what the graph wants.)
        # The images packed here are a snapshot of the weights at forward time; backward reuses exactly these.
        fwd, bias_packed, dg = cp.packed(need_dgrad=ctx.needs_input_grad[0])
        if cp.shuffle:
            out = torch.empty(B, 2 * H, 2 * W, cp.cout_p // 4, dtype=torch.bfloat16, device=x.device)
        else:
            out = torch.empty(B, H, W, cp.cout_p, dtype=torch.float32 if cp.out_f32 else torch.bfloat16, device=x.device)
        d = cp.desc(B, H, W)
        if cp.out_f32:      # last layer of a model: keep the fp32 accumulator (plc_conv_fwd_f32)
            _call(x, lib.plc_conv_fwd_f32, "plc_conv_fwd_f32", ctypes.byref(d), _ptr(x), _ptr(fwd), _ptr(bias_packed), _ptr(out))
        else:
            _call(x, lib.plc_conv_fwd, "plc_conv_fwd", ctypes.byref(d), _ptr(x), _ptr(fwd), _ptr(bias_packed), _ptr(out))
        ctx.cp, ctx.dg = cp, dg
        ctx.save_for_backward(x, out)
        ctx.x_needs_grad = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, out = ctx.saved_tensors
        cp = ctx.cp
        B, H, W, _ = x.shape
        dg = ctx.dg
        d = cp.desc(B, H, W)
        dy = dy.contiguous()
        if dy.dtype != torch.bfloat16:                     # fp32-output layer: the gradient enters the bf16 backward here
            dy = dy.to(torch.bfloat16)
        if cp.relu or cp.shuffle:
            dz = torch.empty(B, H, W, cp.cout_p, dtype=torch.bfloat16, device=x.device)
            _call(dy, lib.plc_conv_grad_mask, "plc_conv_grad_mask", ctypes.byref(d), _ptr(out), _ptr(dy), _ptr(dz))
        else:
            dz = dy
        if (not ctx.x_needs_grad) and cp.cin_p == 8 and cp.Cin * cp.k * cp.k <= 32 and not cp.shuffle:
            # narrow input (front-end init_conv: 3 channels): one 32-column im2col row per pixel, then the reduction
            # over pixels is ONE [Cout x 32] column block instead of k*k blocks of mostly zero padding
            col = torch.empty(B, H, W, 32, dtype=torch.bfloat16, device=x.device)
            _call(x, lib.plc_conv_im2col_narrow, "plc_conv_im2col_narrow", ctypes.byref(d), cp.Cin, _ptr(x), _ptr(col))
            d1 = PlcConvDesc(B, H, W, 32, cp.cout_p, 1, 0, 0, int(cp.conv.bias is not None))
            dW1 = torch.zeros(cp.cout_p, 32, 1, 1, dtype=torch.float32, device=x.device)
            img = torch.zeros(lib.plc_conv_wgrad_acc_bytes(ctypes.byref(d1)) // 4, dtype=torch.float32, device=x.device)
            db = torch.zeros(cp.cout_p, dtype=torch.float32, device=x.device) if cp.conv.bias is not None else None
            _call(x, lib.plc_conv_bwd, "plc_conv_bwd", ctypes.byref(d1), _ptr(col), _ptr(dz), None, None, _ptr(img), _ptr(db))
            _call(img, lib.plc_conv_wgrad_unpack, "plc_conv_wgrad_unpack", ctypes.byref(d1), _ptr(img), _ptr(dW1))
            kk = cp.k * cp.k
            gw = dW1[:cp.Cout, :kk * cp.Cin, 0, 0].reshape(cp.Cout, kk, cp.Cin).permute(0, 2, 1)
            gw = gw.reshape(cp.Cout, cp.Cin, cp.k, cp.k).to(cp.conv.weight.dtype)
            gb = None if db is None else db[:cp.Cout].to(cp.conv.bias.dtype)
            return None, gw, gb, None
        dx = torch.empty_like(x) if ctx.x_needs_grad else None
        dW = torch.zeros(cp.cout_p, cp.cin_p, cp.k, cp.k, dtype=torch.float32, device=x.device)
        img = torch.zeros(lib.plc_conv_wgrad_acc_bytes(ctypes.byref(d)) // 4, dtype=torch.float32, device=x.device)
        db = torch.zeros(cp.cout_p, dtype=torch.float32, device=x.device) if cp.conv.bias is not None else None
        _call(x, lib.plc_conv_bwd, "plc_conv_bwd", ctypes.byref(d), _ptr(x), _ptr(dz), _ptr(dg), _ptr(dx), _ptr(img), _ptr(db))
        _call(img, lib.plc_conv_wgrad_unpack, "plc_conv_wgrad_unpack", ctypes.byref(d), _ptr(img), _ptr(dW))
        gw = dW[:cp.Cout, :cp.Cin].to(cp.conv.weight.dtype)
        gb = None if db is None else db[:cp.Cout].to(cp.conv.bias.dtype)
        return dx, gw, gb, None


def conv2d_same(x: Tensor, cp: ConvParams) -> Tensor:
    """x [B,H,W,cin_p] bf16 -> [B,H,W,cout_p] (or [B,2H,2W,cout_p/4] with PixelShuffle); padded channels are zero.
    bf16 output, fp32 when ``cp.out_f32``."""
    return _ConvFn.apply(x, cp.conv.weight, cp.conv.bias, cp)


def frames_to_nhwc(frames: Tensor, c_pad: int, out: Optional[Tensor] = None) -> Tensor:
    """frames [B,T,Cf,H,W] fp32 -> [T*B,H,W,c_pad] bf16 (T-major) with the coordconv planes appended
    (coordconv.py:3-10); the input of init_conv on the tensor-core conv."""
    lib = _lib.load()
    B, T, Cf, H, W = frames.shape
    _require_cuda(frames, "frames")
    if frames.dtype != torch.float32:
        raise RuntimeError("frames must be float32")
    if out is None:
        out = torch.empty(T * B, H, W, c_pad, dtype=torch.bfloat16, device=frames.device)
    _call(frames, lib.plc_frames_to_nhwc, "plc_frames_to_nhwc", _ptr(frames), B, T, Cf, H, W, c_pad, _ptr(out))
    return out


def frontend_tc_supported(cf: int, c: int, c_stride: int) -> bool:
    return bool(_lib.load().plc_frontend_tc_supported(cf, c, c_stride))


def frontend_tc(frames: Tensor, weight: Tensor, bias: Optional[Tensor], out: Tensor) -> Tensor:
    """relu(init_conv(add_coord_channels(frame))) for all T frames (generator.py:166-168) in one tensor-core launch with
    in-kernel im2col: frames [B,T,Cf,H,W] fp32, weight [64,Cf+2,3,3] fp32 -> out [T*B,H,W,64] bf16 (T-major)."""
    lib = _lib.load()
    B, T, Cf, H, W = frames.shape
    _require_cuda(frames, "frames")
    _require_cuda(out, "out")
    if frames.dtype != torch.float32 or out.dtype != torch.bfloat16 or tuple(out.shape) != (T * B, H, W, 64):
        raise RuntimeError("frontend_tc: frames must be fp32 [B,T,Cf,H,W], out bf16 [T*B,H,W,64]")
    w = weight.detach().to(torch.float32).contiguous()
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    _call(frames, lib.plc_frontend_tc_fwd, "plc_frontend_tc_fwd", _ptr(frames), B, T, Cf, H, W, _ptr(w), _ptr(b), 64, _ptr(out))
    return out


class _FrontendTcFn(torch.autograd.Function):
    """Training form of :func:`frontend_tc`: forward = the im2col tensor-core front-end kernel straight from the fp32
    frames (no NHWC / coordinate-plane tensor, 3x less time than the generic conv on an 8-channel padded input);
    backward = ReLU mask + the narrow-input weight gradient (plc_conv_im2col_narrow + one [64 x 32] column block) on
    the 8-channel NHWC frames, which are only built here.  The frames take no gradient."""

    @staticmethod
    def forward(ctx, frames, weight, bias, cp: ConvParams):
        B, T, Cf, H, W = frames.shape
        out = torch.empty(T * B, H, W, 64, dtype=torch.bfloat16, device=frames.device)
        frontend_tc(frames, weight, bias, out)
        ctx.cp = cp
        ctx.save_for_backward(frames, out)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        frames, out = ctx.saved_tensors
        cp = ctx.cp
        B, T, Cf, H, W = frames.shape
        n = T * B
        d = cp.desc(n, H, W)
        dy = dy.contiguous()
        dz = torch.empty(n, H, W, cp.cout_p, dtype=torch.bfloat16, device=frames.device)
        _call(dy, lib.plc_conv_grad_mask, "plc_conv_grad_mask", ctypes.byref(d), _ptr(out), _ptr(dy), _ptr(dz))
        x = frames_to_nhwc(frames, cp.cin_p)                                      # + coord planes, bf16 [T*B,H,W,8]
        col = torch.empty(n, H, W, 32, dtype=torch.bfloat16, device=frames.device)
        _call(x, lib.plc_conv_im2col_narrow, "plc_conv_im2col_narrow", ctypes.byref(d), cp.Cin, _ptr(x), _ptr(col))
        d1 = PlcConvDesc(n, H, W, 32, cp.cout_p, 1, 0, 0, int(cp.conv.bias is not None))
        dW1 = torch.zeros(cp.cout_p, 32, 1, 1, dtype=torch.float32, device=frames.device)
        img = torch.zeros(lib.plc_conv_wgrad_acc_bytes(ctypes.byref(d1)) // 4, dtype=torch.float32, device=frames.device)
        db = torch.zeros(cp.cout_p, dtype=torch.float32, device=frames.device) if cp.conv.bias is not None else None
        _call(x, lib.plc_conv_bwd, "plc_conv_bwd", ctypes.byref(d1), _ptr(col), _ptr(dz), None, None, _ptr(img), _ptr(db))
        _call(img, lib.plc_conv_wgrad_unpack, "plc_conv_wgrad_unpack", ctypes.byref(d1), _ptr(img), _ptr(dW1))
        kk = cp.k * cp.k
        gw = dW1[:cp.Cout, :kk * cp.Cin, 0, 0].reshape(cp.Cout, kk, cp.Cin).permute(0, 2, 1)
        gw = gw.reshape(cp.Cout, cp.Cin, cp.k, cp.k).to(cp.conv.weight.dtype)
        gb = None if db is None else db[:cp.Cout].to(cp.conv.bias.dtype)
        return None, gw, gb, None


def frontend_tc_train(frames: Tensor, cp: ConvParams) -> Tensor:
    """Differentiable (w.r.t. the conv parameters) fused front-end: frames [B,T,Cf,H,W] fp32 -> [T*B,H,W,64] bf16."""
    return _FrontendTcFn.apply(frames, cp.conv.weight, cp.conv.bias, cp)


def conv2d_same_into(x: Tensor, cp: ConvParams, out: Tensor) -> Tensor:
    """Inference-only conv into a preallocated buffer (no autograd node)."""
    lib = _lib.load()
    B, H, W, _ = x.shape
    fwd, bias_packed, _ = cp.packed(need_dgrad=False)
    d = cp.desc(B, H, W)
    _call(x, lib.plc_conv_fwd, "plc_conv_fwd", ctypes.byref(d), _ptr(x), _ptr(fwd), _ptr(bias_packed), _ptr(out))
    return out


# --------------------------------------------------------------------------------- strided 2-D / 3-D conv (bf16)
from ._lib import PlcConvNdDesc  # noqa: E402


class ConvNdParams:
    """Kernel-ready images of one (possibly strided, possibly 3-D) conv layer, cached per parameter version.

    Wraps an ``nn.Conv2d`` (weight [Cout,Cin,k,k]) or ``nn.Conv3d`` (weight [Cout,Cin,kt,k,k]) PARAMETER HOLDER with
    odd kernel sizes, padding k//2 and strides 1 or 2.  Channel counts are zero-padded to multiples of 8.
    act: 0 none, 1 ReLU, 2 LeakyReLU(slope)."""

    def __init__(self, conv: torch.nn.Module, act: int = 0, slope: float = 0.2):
        self.conv, self.act, self.slope = conv, act, float(slope)
        w = conv.weight
        self.is3d = w.dim() == 5
        self.Cout, self.Cin = w.shape[0], w.shape[1]
        self.kt = w.shape[2] if self.is3d else 1
        self.k = w.shape[-1]
        st = conv.stride
        self.stride_t = st[0] if self.is3d else 1
        self.stride = st[-1]
        if w.shape[-1] != w.shape[-2] or st[-1] != st[-2] or any(p != kk // 2 for p, kk in zip(conv.padding, w.shape[2:])):
            raise ValueError("ConvNdParams: square spatial kernels / strides with padding k//2 only")
        self.cin_p, self.cout_p = _rup(self.Cin, 8), _rup(self.Cout, 8)
        self._cache = None

    def desc(self, B, T, H, W) -> PlcConvNdDesc:
        return PlcConvNdDesc(B, T, H, W, self.cin_p, self.cout_p, self.kt, self.k, self.stride_t, self.stride, self.act,
                             self.slope, int(self.conv.bias is not None))

    def packed(self, need_dgrad: bool):
        w, b = self.conv.weight, self.conv.bias
        key = (_lib.weight_generation(), w.data_ptr(), w._version,
               None if b is None else (b.data_ptr(), b._version), str(w.device))
        pc = self._cache
        if pc is not None and pc[0] == key and (pc[3] is not None or not need_dgrad):
            return pc[1], pc[2], pc[3]
        lib = _lib.load()
        wp = torch.zeros(self.cout_p, self.cin_p, self.kt, self.k, self.k, device=w.device, dtype=torch.float32)
        wp[:self.Cout, :self.Cin] = w.detach().to(torch.float32).reshape(self.Cout, self.Cin, self.kt, self.k, self.k)
        bp = None
        if b is not None:
            bp = torch.zeros(self.cout_p, device=w.device, dtype=torch.float32)
            bp[:self.Cout] = b.detach().to(torch.float32)
        d = self.desc(1, 1, 1, 1)
        fwd = torch.empty(lib.plc_convnd_packed_weight_bytes(ctypes.byref(d), PLC_PACK_FWD), dtype=torch.uint8,
                          device=w.device)
        bias_packed = torch.zeros(self.cout_p, device=w.device, dtype=torch.float32)
        _call(wp, lib.plc_convnd_pack_weight, "plc_convnd_pack_weight", ctypes.byref(d), PLC_PACK_FWD, _ptr(wp), _ptr(bp),
              _ptr(fwd), _ptr(bias_packed))
        dg = None
        if need_dgrad:
            dg = torch.empty(lib.plc_convnd_packed_weight_bytes(ctypes.byref(d), PLC_PACK_DGRAD), dtype=torch.uint8,
                             device=w.device)
            _call(wp, lib.plc_convnd_pack_weight, "plc_convnd_pack_weight", ctypes.byref(d), PLC_PACK_DGRAD, _ptr(wp),
                  None, _ptr(dg), None)
        self._cache = (key, fwd, bias_packed, dg)
        return fwd, bias_packed, dg


class _ConvNdFn(torch.autograd.Function):
    """(strided / 3-D) conv + bias + activation on NHWC bf16 tensors [B,T,H,W,C] through plc_convnd_fwd / plc_convnd_bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias, cp: ConvNdParams):
        lib = _lib.load()
        B, T, H, W, C = x.shape
        if C != cp.cin_p or x.dtype != torch.bfloat16 or not x.is_contiguous():
            raise RuntimeError(f"convnd input must be contiguous bf16 [B,T,H,W,{cp.cin_p}], got {x.dtype} {tuple(x.shape)}")
        if not cp.is3d and (cp.kt != 1 or cp.stride_t != 1):
            raise RuntimeError("2-D layer with a time kernel")
        need_dx = ctx.needs_input_grad[0]
        fwd, bias_packed, dg = cp.packed(need_dgrad=need_dx)       # snapshot of the weights this forward ran with
        d = cp.desc(B, T, H, W)
        to, ho, wo = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.plc_convnd_out_shape(ctypes.byref(d), ctypes.byref(to), ctypes.byref(ho), ctypes.byref(wo)),
                   "plc_convnd_out_shape")
        out = torch.empty(B, to.value, ho.value, wo.value, cp.cout_p, dtype=torch.bfloat16, device=x.device)
        _call(x, lib.plc_convnd_fwd, "plc_convnd_fwd", ctypes.byref(d), _ptr(x), _ptr(fwd), _ptr(bias_packed), _ptr(out))
        ctx.cp, ctx.dg, ctx.d = cp, dg, d
        ctx.save_for_backward(x, out)
        ctx.need_dx, ctx.need_dw = need_dx, ctx.needs_input_grad[1]
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, out = ctx.saved_tensors
        cp, d = ctx.cp, ctx.d
        dy = dy.contiguous()
        if cp.act != 0:
            dz = torch.empty_like(out)
            _call(dy, lib.plc_convnd_grad_mask, "plc_convnd_grad_mask", ctypes.byref(d), _ptr(out), _ptr(dy), _ptr(dz))
        else:
            dz = dy
        dx = torch.empty_like(x) if ctx.need_dx else None
        gw = gb = None
        img = db = None
        if ctx.need_dw:
            img = torch.zeros(lib.plc_convnd_wgrad_acc_bytes(ctypes.byref(d)) // 4, dtype=torch.float32, device=x.device)
            db = torch.zeros(cp.cout_p, dtype=torch.float32, device=x.device) if cp.conv.bias is not None else None
        _call(x, lib.plc_convnd_bwd, "plc_convnd_bwd", ctypes.byref(d), _ptr(x), _ptr(dz), _ptr(ctx.dg), _ptr(dx), _ptr(img),
              _ptr(db))
        if ctx.need_dw:
            dW = torch.zeros(cp.cout_p, cp.cin_p, cp.kt, cp.k, cp.k, dtype=torch.float32, device=x.device)
            _call(img, lib.plc_convnd_wgrad_unpack, "plc_convnd_wgrad_unpack", ctypes.byref(d), _ptr(img), _ptr(dW))
            gw = dW[:cp.Cout, :cp.Cin].reshape(cp.conv.weight.shape).to(cp.conv.weight.dtype)
            gb = None if db is None else db[:cp.Cout].to(cp.conv.bias.dtype)
        return dx, gw, gb, None


def convnd(x: Tensor, cp: ConvNdParams) -> Tensor:
    """x [B,T,H,W,cin_p] bf16 (T = 1 for images) -> [B,To,Ho,Wo,cout_p] bf16; padded channels are zero."""
    return _ConvNdFn.apply(x, cp.conv.weight, cp.conv.bias, cp)


# --------------------------------------------------------------------------------- frame-level first layer (fp32 frames in)
from ._lib import PlcFrameConvDesc  # noqa: E402


def frameconv_supported(conv: torch.nn.Module) -> bool:
    """True if `conv` is a layer plc_frameconv_* runs: Conv2d(Cf <= 4, Cout in {8..256, power of two}, 3, stride 1|2, padding 1)."""
    w = conv.weight
    g = w.shape[0] // 8
    return (w.dim() == 4 and tuple(w.shape[2:]) == (3, 3) and w.shape[1] <= 4 and w.shape[0] % 8 == 0 and 1 <= g <= 32
            and g & (g - 1) == 0 and tuple(conv.padding) == (1, 1) and conv.stride[0] == conv.stride[1]
            and conv.stride[0] in (1, 2))


class _FrameConvFn(torch.autograd.Function):
    """act(conv3x3(frames)) from fp32 frames [N,Cf,H,W] to NHWC bf16 [N,Ho,Wo,Cout] (plc_frameconv_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, frames, weight, bias, stride: int, act: int, slope: float):
        lib = _lib.load()
        _require_cuda(frames, "frames")
        if frames.dtype != torch.float32 or frames.dim() != 4 or not frames.is_contiguous():
            raise RuntimeError(f"frameconv: frames must be contiguous fp32 [N,Cf,H,W], got {frames.dtype} {tuple(frames.shape)}")
        N, Cf, H, W = frames.shape
        d = PlcFrameConvDesc(N, Cf, H, W, weight.shape[0], stride, act, float(slope), int(bias is not None))
        ho, wo = ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.plc_frameconv_out_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)), "plc_frameconv_out_shape")
        w = weight.detach().to(torch.float32).contiguous()       # snapshot of the weights this forward ran with
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        out = torch.empty(N, ho.value, wo.value, weight.shape[0], dtype=torch.bfloat16, device=frames.device)
        _call(frames, lib.plc_frameconv_fwd, "plc_frameconv_fwd", ctypes.byref(d), _ptr(frames), _ptr(w), _ptr(b), _ptr(out))
        ctx.d, ctx.w, ctx.has_bias = d, w, bias is not None
        ctx.save_for_backward(frames, out)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        frames, out = ctx.saved_tensors
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dy = dy.contiguous()
        dx = torch.empty_like(frames) if need_dx else None
        dW = torch.zeros_like(ctx.w) if need_dw else None
        db = torch.zeros(ctx.w.shape[0], dtype=torch.float32, device=frames.device) if (need_dw and ctx.has_bias) else None
        _call(frames, lib.plc_frameconv_bwd, "plc_frameconv_bwd", ctypes.byref(ctx.d), _ptr(frames), _ptr(ctx.w), _ptr(out),
              _ptr(dy), _ptr(dx), _ptr(dW), _ptr(db))
        return dx, dW, db, None, None, None


def frameconv(frames: Tensor, conv: torch.nn.Module, act: int = 0, slope: float = 0.2) -> Tensor:
    """frames [N,Cf,H,W] fp32 -> act(conv(frames)) as NHWC bf16 [N,Ho,Wo,Cout]; `conv` is the parameter holder."""
    return _FrameConvFn.apply(frames, conv.weight, conv.bias, conv.stride[0], act, slope)
