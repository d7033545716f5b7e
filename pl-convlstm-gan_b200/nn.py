"""Host-side mirror of the reference nn.Module API for the ConvLSTM recurrence.

* :class:`ConvLSTMCell` -- drop-in for ``src/models/convlstm.py:4-28``: same constructor
  ``(input_dim, hidden_dim, kernel_size=3, bias=True)``, same ``hidden_dim`` attribute, same
  ``conv.weight [4Ch, Cin+Ch, k, k]`` / ``conv.bias [4Ch]`` parameters (hence ``state_dict`` keys and the
  default-init RNG stream), same ``forward(x, h_cur, c_cur) -> (h_next, c_next)`` on logical
  ``[B, C, H, W]`` tensors.  ``forward(x, (h, c))`` is accepted too (north_star spelling).
  The arithmetic runs in libplc.so (no cuDNN / ATen conv, no CPU path).
* :class:`ConvLSTMStack` -- the stacked-cell T-loop of ``src/models/generator.py:156-171`` (zero initial
  state, layer l consumes h of layer l-1 at the same step), generalised to L layers and kept in the NHWC
  bf16 working layout between steps.
* :class:`EncoderForecaster` -- north_star extension with no reference counterpart (spec:
  oracle/convlstm_oracle.py:encoder_forecaster_forward).
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import functional as F
from ._lib import PLC_MODE_BF16_TC, PLC_MODE_FP32

Tensor = torch.Tensor

_MODES = {"bf16": PLC_MODE_BF16_TC, "fp32": PLC_MODE_FP32}


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


class _CellStepFn(torch.autograd.Function):
    """One cell step on working-layout (NHWC) tensors, differentiable (backward = plc_cell_bwd)."""

    @staticmethod
    def forward(ctx, x, h, c, weight, bias, cell):
        # torch.is_grad_enabled() is always False in here; what the graph needs is in ctx.needs_input_grad.  The packed
        # images are a SNAPSHOT of the weights at forward time and backward recomputes the gates from that snapshot, so
        # an optimizer step between forward and backward (GAN loops) cannot desynchronise the gradient.
        pw = cell._packed(need_dgrad=any(ctx.needs_input_grad))
        h2, c2 = F.cell_forward(x, h, c, pw)
        ctx.cell = cell
        ctx.pw = pw
        ctx.has_x = x is not None
        ctx.x_needs_grad = x is not None and x.requires_grad
        ctx.save_for_backward(x, h, c)
        return h2, c2

    @staticmethod
    def backward(ctx, dh, dc):
        x, h, c = ctx.saved_tensors
        cell, pw = ctx.cell, ctx.pw
        B, H, W, Ch = h.shape
        if dh is None:
            dh = torch.zeros_like(h)
        dh = dh.contiguous()
        dc = None if dc is None else dc.contiguous()
        dW = torch.zeros(4 * Ch, pw.Cin + Ch, pw.k, pw.k, device=h.device, dtype=torch.float32)
        db = torch.zeros(4 * Ch, device=h.device, dtype=torch.float32) if pw.bias is not None else None
        dx, dh_prev, dc_prev = F.cell_backward(x if ctx.has_x else None, h, c, pw, dh, None, dc, dW, db,
                                               need_dx=ctx.x_needs_grad)
        cin = cell.input_dim
        if pw.Cin != cin:  # drop the zero-padded x channels
            dW = torch.cat([dW[:, :cin], dW[:, pw.Cin:]], dim=1)
        gw = dW.to(cell.conv.weight.dtype)
        gb = None if db is None else db.to(cell.conv.bias.dtype)
        return (dx if ctx.x_needs_grad else None), dh_prev, dc_prev, gw, gb, None


class ConvLSTMCell(nn.Module):
    """Drop-in for the reference ``ConvLSTMCell`` (convlstm.py:4-28), computed by libplc.so.

    mode: "bf16" (tcgen05 tensor cores, fp32 accumulate/state; <=1e-2 rel. vs reference) or
          "fp32" (validation mode; <=1e-5).
    """

    def __init__(self, input_dim: int, hidden_dim: int, kernel_size: int = 3, bias: bool = True,
                 mode: str = "bf16"):
        super().__init__()
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {list(_MODES)}")
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim                     # convlstm.py:7
        self.kernel_size = kernel_size
        self.mode = mode
        # Parameter holder only -- identical to convlstm.py:8-14 so that state_dict keys, shapes and the
        # default init are the reference's.  It is never *called*.
        if input_dim + hidden_dim > 0:
            self.conv = nn.Conv2d(input_dim + hidden_dim, 4 * hidden_dim, kernel_size,
                                  padding=kernel_size // 2, bias=bias)
        self._pack_cache = None

    # -- packed-weight cache, invalidated when the parameters change (optimizer step, load_state_dict, .to())
    def _packed(self, need_dgrad: bool) -> F.PackedWeights:
        w, b = self.conv.weight, self.conv.bias
        key = (_lib.weight_generation(), w.data_ptr(), w._version,
               None if b is None else (b.data_ptr(), b._version), self.mode, str(w.device))
        pc = self._pack_cache
        if pc is not None and pc[0] == key and (pc[1].dgrad is not None or not need_dgrad):
            return pc[1]
        mode = _MODES[self.mode]
        cin_pad = _pad8(self.input_dim) if mode == PLC_MODE_BF16_TC else self.input_dim
        pw = F.pack_weights(w, b, self.input_dim, self.hidden_dim, self.kernel_size, mode,
                            with_dgrad=need_dgrad, cin_pad=cin_pad)
        self._pack_cache = (key, pw)
        return pw

    @property
    def working_cin(self) -> int:
        return _pad8(self.input_dim) if self.mode == "bf16" else self.input_dim

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.mode == "bf16" else torch.float32

    def step_nhwc(self, x: Optional[Tensor], h: Tensor, c: Tensor) -> Tuple[Tensor, Tensor]:
        """Working-layout step: x [B,H,W,working_cin] / h [B,H,W,Ch] in act_dtype, c fp32.  Differentiable."""
        needs_grad = torch.is_grad_enabled() and (
            self.conv.weight.requires_grad or h.requires_grad or c.requires_grad or (x is not None and x.requires_grad))
        if needs_grad:
            return _CellStepFn.apply(x, h, c, self.conv.weight, self.conv.bias, self)
        return F.cell_forward(x, h, c, self._packed(need_dgrad=False))

    def _to_working(self, t: Tensor, channels: int, dtype) -> Tensor:
        # logical [B,C,H,W] (any strides) -> NHWC contiguous [B,H,W,C'] ; differentiable torch ops (plumbing)
        t = t.permute(0, 2, 3, 1)
        if channels != t.shape[-1]:
            t = torch.nn.functional.pad(t, (0, channels - t.shape[-1]))
        return t.to(dtype).contiguous()

    def forward(self, x: Tensor, h_cur, c_cur: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """convlstm.py:16-28.  Logical [B,C,H,W] in, ``(h_next, c_next)`` logical [B,C,H,W] out (dtype of
        ``h_cur`` / ``c_cur`` respectively, channels_last strides).

        Cost of this literal drop-in form: logical NCHW-contiguous fp32 inputs are converted on the way in (3 permute /
        cast copies) and out (2 casts).  Callers that keep the recurrent state in the working layout pay nothing: pass
        ``x`` / ``h_cur`` as bf16 and ``c_cur`` as fp32 tensors with ``torch.channels_last`` strides (what this method
        returns) and every conversion above is a view -- or use :class:`ConvLSTMStack`, which never leaves NHWC."""
        if c_cur is None:
            h_cur, c_cur = h_cur                         # forward(x, (h, c)) spelling
        if not h_cur.is_cuda:
            raise RuntimeError("ConvLSTMCell (plconv) has no CPU path: move the module and inputs to a CUDA device")
        if self.kernel_size % 2 == 0:
            raise RuntimeError("kernel_size must be odd (the reference's padding k//2 breaks even k)")
        out_dtype, c_dtype = h_cur.dtype, c_cur.dtype
        xw = self._to_working(x, self.working_cin, self.act_dtype) if self.input_dim > 0 else None
        hw = self._to_working(h_cur, self.hidden_dim, self.act_dtype)
        cw = self._to_working(c_cur, self.hidden_dim, torch.float32)
        h2, c2 = self.step_nhwc(xw, hw, cw)
        # back to the logical NCHW shape (a channels_last-strided view; no copy beyond the dtype cast)
        return h2.to(out_dtype).permute(0, 3, 1, 2), c2.to(c_dtype).permute(0, 3, 1, 2)


# Saved-gates BPTT (plc_cell_fwd_save / plc_cell_bwd_saved): "auto" keeps the activated gates of a rollout when they fit
# comfortably (<= 20 % of the device's memory and <= half of what is free right now), "on" always (where the shape has
# a saved form), "off" never (recompute, no extra memory).  PLC_SAVE_GATES overrides the default.
SAVE_GATES = os.environ.get("PLC_SAVE_GATES", "auto")
LAST_SAVED_GATES_BYTES = 0      # introspection: bytes the most recent rollout kept (0 = it recomputed)


def _saved_gates_plan(cells, pws, T, B, H, W, dev):
    """bytes per step of every layer's saved-gates buffer (0 = recompute for that layer)."""
    global LAST_SAVED_GATES_BYTES
    mode = SAVE_GATES
    LAST_SAVED_GATES_BYTES = 0
    if mode not in ("auto", "on", "1"):
        return [0] * len(cells)
    per = [F.saved_gates_bytes(B, H, W, pw) for pw in pws]
    total = T * sum(per)
    if mode == "auto" and total > 0:
        free, cap = torch.cuda.mem_get_info(dev)
        free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        if total > 0.2 * cap or total > 0.5 * free:
            return [0] * len(cells)
    LAST_SAVED_GATES_BYTES = total
    return per


# Chunked weight gradient (plc_cell_wgrad): keep dZ of `chunk` consecutive steps of a layer and run ONE wgrad launch over
# [chunk*B] images instead of one per step -- the accumulator flush, prologue and tail of a wgrad launch are fixed costs
# (~12 us) that weigh 1 % at 64 sequences per GPU and ~20 % of the kernel at 8 (cfg3 on 8 GPUs: 11.27 -> 10.83 ms per
# step).  The chunk is sized so that a launch reduces about as many pixels as one step at 64 sequences of 128x128 does
# (2^20): longer reductions are NOT free -- the tensor core's fp32 accumulation error grows with the chain length
# (measured: all T = 10 steps of B = 64 in one launch moves dW by 1e-3 of its max, one step per launch by 1e-4), so the
# chunk never exceeds the chain length the full-size parity tests validate.  PLC_DEFER_WGRAD = auto | off | <steps>.
DEFER_WGRAD = os.environ.get("PLC_DEFER_WGRAD", "auto")
_WGRAD_CHAIN_PIXELS = 1 << 20


def _wgrad_chunk(T, B, H, W):
    """steps per wgrad launch (1 = the per-step form inside plc_cell_bwd)."""
    mode = str(DEFER_WGRAD)
    if mode in ("off", "0", "1") or T < 2:
        return 1
    if mode.isdigit():
        return max(1, min(T, int(mode)))
    return max(1, min(T, _WGRAD_CHAIN_PIXELS // max(1, B * H * W)))


# Layer wavefront (opt-in, PLC_LAYER_STREAMS=1): cell (t, l) depends on (t-1, l) and (t, l-1) only, so layer l's step t and
# layer l-1's step t+1 are independent -- forward and backward -- and each layer's launches can go to their OWN stream
# with event edges between layers (captured into CUDA graphs as parallel branches).  Measured on B200 and NOT a gain
# (cfg3 at 8 sequences per GPU: 11.13 vs 11.03 ms per step; at 64: 79.7 vs 79.4): every kernel here is persistent with
# one CTA per SM and ~200 KB of shared memory, so a second kernel's CTA can only start when the first one's CTA has
# exited -- the idle tail of a launch (the last tile's epilogue, ~5 us) lies INSIDE the CTA's lifetime and cannot be
# covered by another stream, and the launch / prologue gap is already hidden by programmatic dependent launch.  Kept
# for multi-layer stacks with small tiles; default off.
LAYER_STREAMS = os.environ.get("PLC_LAYER_STREAMS", "0") == "1"
_side_streams = {}


def _layer_streams(dev, L):
    """[current stream] + (L-1) side streams of `dev` (created once per device), or None when disabled / L == 1."""
    if not LAYER_STREAMS or L < 2:
        return None
    pool = _side_streams.setdefault(dev.index, [])
    while len(pool) < L - 1:
        pool.append(torch.cuda.Stream(dev))
    return [torch.cuda.current_stream(dev)] + pool[:L - 1]


class _StackRolloutFn(torch.autograd.Function):
    """The whole stacked T-step rollout (generator.py:156-171) as ONE autograd node.

    forward : T x L fused cell steps into preallocated state rings  h[l][0..T], c[l][0..T]  (index 0 = initial state).
    backward: explicit BPTT -- per cell step one plc_cell_bwd with
              dh = gradient from the layer above / the loss, dh2 = recurrent gradient from step t+1 (summed inside the
              kernel), dW / db accumulated in place over all T steps.  No per-step autograd nodes, no per-step
              allocations, gates recomputed from the saved (x, h_prev, c_prev).
    Inputs : xs [T,B,H,W,Cin'] or None, then (h0_l, c0_l) per layer, then (weight_l, bias_l) per layer.
    Outputs: top-layer h for every step [T,B,H,W,Ch_L], then final (h_l, c_l) per layer.
    """

    @staticmethod
    def forward(ctx, cells, T, xs, *tensors):
        L = len(cells)
        states, params = tensors[:2 * L], tensors[2 * L:]
        need_grad = any(ctx.needs_input_grad)          # (torch.is_grad_enabled() is always False inside forward)
        pws = [c._packed(need_dgrad=need_grad) for c in cells]   # snapshot of the weights this rollout ran with
        B, H, W, _ = states[0].shape
        dev = states[0].device
        hs, cs = [], []
        for l, cell in enumerate(cells):
            h_ring = torch.empty(T + 1, B, H, W, cell.hidden_dim, device=dev, dtype=cell.act_dtype)
            c_ring = torch.empty(T + 1, B, H, W, cell.hidden_dim, device=dev, dtype=torch.float32)
            h_ring[0].copy_(states[2 * l])
            c_ring[0].copy_(states[2 * l + 1])
            hs.append(h_ring)
            cs.append(c_ring)
        # saved-gates BPTT where memory allows: one buffer per (layer, step), filled by the forward kernel
        sv = [None] * L
        if need_grad:
            per = _saved_gates_plan(cells, pws, T, B, H, W, dev)
            sv = [torch.empty(T, n, device=dev, dtype=torch.uint8) if n else None for n in per]
        streams = _layer_streams(dev, L)
        if streams is None:
            for t in range(T):                                   # generator.py:164
                inp = None if xs is None else xs[t]
                for l in range(L):                               # generator.py:170-171
                    F.cell_forward(inp, hs[l][t], cs[l][t], pws[l], h_out=hs[l][t + 1], c_out=cs[l][t + 1],
                                   saved=None if sv[l] is None else sv[l][t])
                    inp = hs[l][t + 1]
        else:
            # wavefront over per-layer streams: layer l, step t waits for layer l-1, step t (its input h)
            fork = torch.cuda.Event()
            fork.record(streams[0])
            for st in streams[1:]:
                st.wait_event(fork)
            done = [None] * L                                    # event after layer l's most recent step
            for t in range(T):
                for l in range(L):
                    with torch.cuda.stream(streams[l]):
                        if l > 0:
                            streams[l].wait_event(done[l - 1])
                        F.cell_forward((None if xs is None else xs[t]) if l == 0 else hs[l - 1][t + 1], hs[l][t], cs[l][t],
                                       pws[l], h_out=hs[l][t + 1], c_out=cs[l][t + 1],
                                       saved=None if sv[l] is None else sv[l][t])
                        if l < L - 1 or t == T - 1:
                            done[l] = torch.cuda.Event()
                            done[l].record(streams[l])
            for l in range(1, L):                                # join: everything after this node sees all layers done
                streams[0].wait_event(done[l])
        ctx.cells, ctx.T, ctx.pws, ctx.sv = cells, T, pws, sv
        ctx.xs, ctx.hs, ctx.cs = xs, hs, cs
        ctx.x_needs_grad = xs is not None and xs.requires_grad
        ctx.state_needs_grad = [s.requires_grad for s in states]
        outs = [hs[L - 1][1:]]
        for l in range(L):
            outs += [hs[l][T], cs[l][T]]
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_out, *d_final):
        cells, T, xs, hs, cs = ctx.cells, ctx.T, ctx.xs, ctx.hs, ctx.cs
        L = len(cells)
        pws = ctx.pws
        B, H, W, _ = hs[0][0].shape
        dev = hs[0].device
        dW_img = [F.wgrad_accumulator(B, H, W, pw, dev) for pw in pws]     # accumulated over all T steps
        db = [torch.zeros(4 * c.hidden_dim, device=dev) if pw.bias is not None else None for c, pw in zip(cells, pws)]
        # dZ: one workspace per layer reused by every step, or -- deferred wgrad -- a ring over all T steps
        ws_bytes = [F.bwd_workspace_bytes(B, H, W, pw) for pw in pws]
        chunk = _wgrad_chunk(T, B, H, W)                     # steps per weight-gradient launch (see DEFER_WGRAD)
        ws = [torch.empty(chunk, n, dtype=torch.uint8, device=dev) for n in ws_bytes]
        # recurrent carries: dh ping-pong (read as dh2 while the next dh_prev is written), dc in place
        dh_buf = [[torch.empty_like(hs[l][0]) for _ in range(2)] for l in range(L)]
        dc_buf = [torch.empty_like(cs[l][0]) for l in range(L)]
        # dx of layer l feeds layer l-1 at the same step: two buffers (step parity), so that with per-layer streams layer l
        # can start step t-1 while layer l-1 still reads step t's
        dx_buf = [[torch.empty(B, H, W, pws[l].Cin, device=dev, dtype=cells[l].act_dtype) for _ in range(2)]
                  if (pws[l].Cin and l > 0) else None for l in range(L)]
        dh_carry = [None] * L
        dc_carry = [None] * L
        for l in range(L):
            dhT, dcT = d_final[2 * l], d_final[2 * l + 1]
            if dhT is not None:
                dh_carry[l] = dhT.contiguous()
            if dcT is not None:
                dc_buf[l].copy_(dcT)
                dc_carry[l] = dc_buf[l]
        dxs = torch.empty_like(xs) if ctx.x_needs_grad else None
        zero_top = {}
        flip = [0] * L
        wgrads = [None] * L

        def chunk_wgrad(l, t0):
            """ONE weight-gradient launch over steps t0 .. t0+n-1 (their dZ sits in ws[l][0:n], in step order): the state
            rings' slices ARE x and h_prev of those steps."""
            n = min(chunk, T - t0)
            if l == 0:
                x_all = None if xs is None else xs[t0:t0 + n].reshape(n * B, H, W, xs.shape[-1])
            else:
                x_all = hs[l - 1][t0 + 1:t0 + n + 1].reshape(n * B, H, W, hs[l - 1].shape[-1])
            F.cell_wgrad(x_all, hs[l][t0:t0 + n].reshape(n * B, H, W, cells[l].hidden_dim), ws[l][0:n], pws[l],
                         dW_img[l], db[l])

        def finish_layer(l):
            """Layer l's last BPTT step (t = 0) has been queued: convert its accumulator to the reference layout and, if
            a gradient sink is attached (GradReducer under data parallelism), hand dW / db over NOW so that the bucket's
            all-reduce overlaps the remaining BPTT steps of the layers below instead of waiting for the whole node."""
            cell = cells[l]
            g = torch.zeros(4 * cell.hidden_dim, pws[l].Cin + cell.hidden_dim, pws[l].k, pws[l].k, device=dev)
            F.wgrad_unpack(dW_img[l], pws[l], g)             # one layout conversion per backward pass
            if pws[l].Cin != cell.input_dim:                 # drop zero-padded x channels
                g = torch.cat([g[:, :cell.input_dim], g[:, pws[l].Cin:]], dim=1)
            gw = g.to(cell.conv.weight.dtype)
            gb = None if cell.conv.bias is None else db[l].to(cell.conv.bias.dtype)
            sink = getattr(cell, "_grad_sink", None)
            if sink is not None and sink(cell, gw, gb):
                wgrads[l] = (None, None)
            else:
                wgrads[l] = (gw, gb)

        if d_out is None:                                    # every layer may need it: allocate before the streams fork
            zero_top = {}
            for l in range(L):
                if hs[l][0].shape not in zero_top:
                    zero_top[hs[l][0].shape] = torch.zeros_like(hs[l][0])
        # BPTT wavefront over per-layer streams (see LAYER_STREAMS): layer l, step t waits for layer l+1, step t (its dh
        # from above) and -- before it overwrites the dx buffer of step t's parity -- for layer l-1's step t+1
        streams = _layer_streams(dev, L)
        done = [None] * L
        if streams is not None:
            fork = torch.cuda.Event()
            fork.record(streams[0])
            for st in streams[1:]:
                st.wait_event(fork)
        for t in reversed(range(T)):
            d_above = None if d_out is None else d_out[t]
            for l in reversed(range(L)):
                with (torch.cuda.stream(streams[l]) if streams is not None else contextlib.nullcontext()):
                    if streams is not None:
                        if l < L - 1 and done[l + 1] is not None:
                            streams[l].wait_event(done[l + 1])       # dx of the layer above, this step
                        if l > 0 and done[l - 1] is not None:
                            streams[l].wait_event(done[l - 1])       # the layer below has consumed step t+1's dx
                    x_in = (xs[t] if xs is not None else None) if l == 0 else hs[l - 1][t + 1]
                    dh, dh2 = d_above, dh_carry[l]
                    if dh is None:
                        dh, dh2 = dh2, None
                    if dh is None:
                        dh = zero_top[hs[l][0].shape]
                    if not dh.is_contiguous():
                        dh = dh.contiguous()
                    need_dx = pws[l].Cin > 0 and (l > 0 or ctx.x_needs_grad)
                    out_dx = None
                    if need_dx:
                        out_dx = dxs[t] if (l == 0) else dx_buf[l][t & 1]
                    dst = dh_buf[l][flip[l]]
                    F.cell_backward_acc(x_in, hs[l][t], cs[l][t], pws[l], dh, dh2, dc_carry[l],
                                        dW_img[l] if chunk == 1 else None, db[l] if chunk == 1 else None,
                                        need_dx=need_dx, workspace=ws[l][t % chunk], dx=out_dx, dh_prev=dst,
                                        dc_prev=dc_buf[l], saved=None if ctx.sv[l] is None else ctx.sv[l][t])
                    if chunk > 1 and t % chunk == 0:
                        chunk_wgrad(l, t)
                    dh_carry[l], dc_carry[l] = dst, dc_buf[l]
                    flip[l] ^= 1
                    d_above = out_dx if l > 0 else None
                    if t == 0:
                        finish_layer(l)
                    if streams is not None:
                        done[l] = torch.cuda.Event()
                        done[l].record(streams[l])
        if streams is not None:
            for l in range(1, L):                            # join: the gradients below are consumed on the node's stream
                if done[l] is not None:
                    streams[0].wait_event(done[l])
        grads = [None, None, dxs]
        for l in range(L):
            grads.append(dh_carry[l] if ctx.state_needs_grad[2 * l] else None)
            grads.append(dc_carry[l] if ctx.state_needs_grad[2 * l + 1] else None)
        for l in range(L):
            if wgrads[l] is None:                            # T == 0: nothing ran
                finish_layer(l)
            grads += list(wgrads[l])
        return tuple(grads)


class ConvLSTMStack(nn.Module):
    """L stacked cells run over T steps (generator.py:156-171 generalised).

    ``cells[0] = ConvLSTMCell(input_dim, hidden_dims[0])``, ``cells[l] = ConvLSTMCell(hidden_dims[l-1],
    hidden_dims[l])`` -- for ``hidden_dims=[a, b]`` and ``input_dim=a`` this is exactly the reference's
    ``cell1``/``cell2`` wiring (generator.py:57-58).
    """

    def __init__(self, input_dim: int, hidden_dims: Sequence[int], kernel_size: int = 3, bias: bool = True,
                 mode: str = "bf16"):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dims = list(hidden_dims)
        dims = [input_dim] + self.hidden_dims
        self.cells = nn.ModuleList(
            [ConvLSTMCell(dims[l], dims[l + 1], kernel_size, bias, mode) for l in range(len(self.hidden_dims))])
        self.mode = mode

    def zero_state(self, B: int, H: int, W: int, device) -> List[Tuple[Tensor, Tensor]]:
        """generator.py:156-160: zeros, fp32 cell state (bf16 h in the working layout)."""
        st = []
        for cell in self.cells:
            st.append((torch.zeros(B, H, W, cell.hidden_dim, device=device, dtype=cell.act_dtype),
                       torch.zeros(B, H, W, cell.hidden_dim, device=device, dtype=torch.float32)))
        return st

    def run_nhwc(self, x_steps: Optional[Sequence[Tensor]], state=None, steps: Optional[int] = None):
        """x_steps: per-step working-layout inputs [B,H,W,working_cin] (or None with ``steps`` for an
        input-less first layer).  Returns (list of top-layer h per step, final state)."""
        T = steps if x_steps is None else len(x_steps)
        if state is None:
            B, H, W, _ = x_steps[0].shape
            state = self.zero_state(B, H, W, x_steps[0].device)
        state = list(state)
        outs = []
        for t in range(T):                                # generator.py:164
            inp = None if x_steps is None else x_steps[t]
            for l, cell in enumerate(self.cells):         # generator.py:170-171
                h, c = state[l]
                h, c = cell.step_nhwc(inp, h, c)
                state[l] = (h, c)
                inp = h
            outs.append(inp)
        return outs, state

    def run_seq(self, xs: Optional[Tensor], state=None, steps: Optional[int] = None):
        """Fused rollout: xs [T,B,H,W,working_cin] (one contiguous tensor) or None with ``steps``.
        Returns (top-layer h for every step [T,B,H,W,Ch_L], final state); one autograd node for the whole rollout."""
        T = steps if xs is None else xs.shape[0]
        if state is None:
            _, B, H, W, _ = xs.shape
            state = self.zero_state(B, H, W, xs.device)
        flat_state = [t for hc in state for t in hc]
        params = []
        for c in self.cells:
            params += [c.conv.weight, c.conv.bias]
        outs = _StackRolloutFn.apply(list(self.cells), T, xs, *flat_state, *params)
        final = [(outs[1 + 2 * l], outs[2 + 2 * l]) for l in range(len(self.cells))]
        return outs[0], final

    def forward(self, x_seq: Tensor, state=None):
        """x_seq: logical [B, T, C, H, W] (fp32, reference layout).  Returns the top layer's h for every
        step as logical [B, T, Ch_L, H, W] fp32, and the final working-layout state."""
        if not x_seq.is_cuda:
            raise RuntimeError("ConvLSTMStack (plconv) has no CPU path")
        B, T, C, H, W = x_seq.shape
        c0 = self.cells[0]
        # one layout conversion for the whole sequence: [B,T,C,H,W] -> T x [B,H,W,C']
        xs = x_seq.permute(1, 0, 3, 4, 2)
        if c0.working_cin != C:
            xs = torch.nn.functional.pad(xs, (0, c0.working_cin - C))
        xs = xs.to(c0.act_dtype).contiguous()
        out, state = self.run_seq(xs, state)              # [T,B,H,W,Ch]
        return out.to(torch.float32).permute(1, 0, 4, 2, 3), state


class EncoderForecaster(nn.Module):
    """Encoder-forecaster ConvLSTM (north_star extension; no reference counterpart -- parity is against
    oracle.encoder_forecaster_forward only).  The forecaster's first layer has no input tensor."""

    def __init__(self, input_dim: int, hidden_dims: Sequence[int], kernel_size: int = 3, t_out: int = 10,
                 mode: str = "bf16"):
        super().__init__()
        self.t_out = t_out
        self.encoder = ConvLSTMStack(input_dim, hidden_dims, kernel_size, True, mode)
        self.forecaster = ConvLSTMStack(0, hidden_dims, kernel_size, True, mode)

    def forward(self, x_seq: Tensor, t_out: Optional[int] = None) -> Tensor:
        t_out = self.t_out if t_out is None else t_out
        _, state = self.encoder(x_seq)
        out, _ = self.forecaster.run_seq(None, state, steps=t_out)
        return out.to(torch.float32).permute(1, 0, 4, 2, 3)
