"""CPU restatement of the reference ConvLSTM recurrence (TEST INFRASTRUCTURE ONLY).

Follows, line by line:

* ``src/models/convlstm.py:16-28``   -- one cell step (cat, conv, split i/f/o/g,
  sigmoid x3 + tanh, ``c' = f*c + i*g``, ``h' = o*tanh(c')``; returns ``(h', c')``)
* ``src/models/convlstm.py:8-14``    -- conv geometry: stride 1, zero padding ``k//2``,
  weight ``[4Ch, Cin+Ch, k, k]`` (x channels first, then h), bias ``[4Ch]``
* ``src/models/generator.py:156-171`` -- zero initial state, stacked cells, layer
  ``l`` consumes ``h`` of layer ``l-1`` at the same time step
* SURVEY.md section 3.3                -- the BPTT equations autograd applies to the above

Written with the same ATen ops the reference dispatches to (``F.conv2d``,
``sigmoid``, ``tanh``) so that, in fp32 on CPU, it *is* the reference arithmetic,
and in fp64 it serves as "truth".  It is pinned against outputs of the unmodified
reference in ``tests/golden`` (see ``tests/test_oracle.py``).

The backward functions are explicit (no autograd) so that they restate the
algorithm the fused CUDA kernels implement; ``tests/test_oracle.py`` checks them
against the reference's autograd.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def conv2d_same(inp: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """``nn.Conv2d(..., kernel_size=k, padding=k//2)`` of convlstm.py:8-14."""
    k = weight.shape[-1]
    return F.conv2d(inp, weight, bias, stride=1, padding=k // 2)


def cell_forward_gates(x: Optional[Tensor], h: Tensor, c: Tensor, weight: Tensor,
                       bias: Optional[Tensor]):
    """convlstm.py:16-28, also returning the gate activations (i, f, o, g).

    ``x`` may be ``None`` (``Cin == 0``: the forecaster's input-less first layer, a
    new-repo extension); then ``weight`` is ``[4Ch, Ch, k, k]``.
    """
    ch = h.shape[1]
    combined = h if x is None else torch.cat([x, h], dim=1)      # convlstm.py:17
    conv_out = conv2d_same(combined, weight, bias)               # convlstm.py:18
    cc_i, cc_f, cc_o, cc_g = torch.split(conv_out, ch, dim=1)    # convlstm.py:19
    i = torch.sigmoid(cc_i)                                      # convlstm.py:21
    f = torch.sigmoid(cc_f)                                      # convlstm.py:22
    o = torch.sigmoid(cc_o)                                      # convlstm.py:23
    g = torch.tanh(cc_g)                                         # convlstm.py:24
    c_next = f * c + i * g                                       # convlstm.py:26
    h_next = o * torch.tanh(c_next)                              # convlstm.py:27
    return h_next, c_next, (i, f, o, g)


def cell_forward(x, h, c, weight, bias) -> Tuple[Tensor, Tensor]:
    """convlstm.py:16-28.  Returns ``(h_next, c_next)`` in that order (convlstm.py:28)."""
    h_next, c_next, _ = cell_forward_gates(x, h, c, weight, bias)
    return h_next, c_next


def cell_backward(x: Optional[Tensor], h_prev: Tensor, c_prev: Tensor, weight: Tensor,
                  bias: Optional[Tensor], dh: Tensor, dc_next: Tensor):
    """Explicit backward of one cell step.

    Inputs: saved ``(x, h_prev, c_prev)``, upstream ``dh`` (w.r.t. ``h_next``) and
    ``dc_next`` (w.r.t. ``c_next``).  Gates are recomputed, as the CUDA path does.
    Returns ``dict(dx, dh_prev, dc_prev, dW, db)``.
    """
    ch = h_prev.shape[1]
    cin = 0 if x is None else x.shape[1]
    k = weight.shape[-1]
    _, c_next, (i, f, o, g) = cell_forward_gates(x, h_prev, c_prev, weight, bias)
    tc = torch.tanh(c_next)
    do = dh * tc
    dc = dc_next + dh * o * (1.0 - tc * tc)
    di = dc * g
    df = dc * c_prev
    dg = dc * i
    dc_prev = dc * f
    dz = torch.cat([di * i * (1.0 - i), df * f * (1.0 - f),
                    do * o * (1.0 - o), dg * (1.0 - g * g)], dim=1)
    combined = h_prev if x is None else torch.cat([x, h_prev], dim=1)
    # dgrad: transpose conv of dz with W (stride 1, "same" padding)
    d_combined = F.conv_transpose2d(dz, weight, None, stride=1, padding=k // 2)
    # wgrad: correlate combined with dz, reduce over batch and pixels
    b = combined.shape[0]
    # conv2d trick: treat batch as channels  -> [Cin+Ch, 4Ch, k, k] -> transpose
    dw = F.conv2d(combined.transpose(0, 1), dz.transpose(0, 1), None, stride=1,
                  padding=k // 2).transpose(0, 1).contiguous()
    assert dw.shape == weight.shape and b == dz.shape[0]
    db = dz.sum(dim=(0, 2, 3))
    dx = d_combined[:, :cin] if cin else None
    dh_prev = d_combined[:, cin:]
    return dict(dx=dx, dh_prev=dh_prev, dc_prev=dc_prev, dW=dw, db=db, dz=dz)


def stack_forward(x_seq: Optional[Tensor], weights: Sequence[Tensor],
                  biases: Sequence[Optional[Tensor]],
                  state: Optional[List[Tuple[Tensor, Tensor]]] = None,
                  steps: Optional[int] = None, return_all: bool = False):
    """Stacked-cell recurrence of generator.py:156-171, generalised to L layers.

    ``x_seq`` is ``[B, T, C, H, W]`` (or ``None`` with ``steps`` given, for an
    input-less first layer).  Zero initial state when ``state is None``
    (generator.py:156-160).  Returns ``(h_top_seq [B,T,Ch_L,H,W], final_state)``
    and, with ``return_all``, every layer's per-step ``(h, c)``.
    """
    n_layers = len(weights)
    t_steps = steps if x_seq is None else x_seq.shape[1]
    ref = weights[0]
    if state is None:
        assert x_seq is not None
        b, _, _, hh, ww = x_seq.shape
        state = []
        for w in weights:
            ch = w.shape[0] // 4
            z = torch.zeros(b, ch, hh, ww, dtype=ref.dtype)       # generator.py:156-160
            state.append((z, z.clone()))
    state = list(state)
    outs, trace = [], []
    for t in range(t_steps):                                      # generator.py:164
        inp = None if x_seq is None else x_seq[:, t]
        step_trace = []
        for l in range(n_layers):                                 # generator.py:170-171
            h, c = state[l]
            h, c = cell_forward(inp, h, c, weights[l], biases[l])
            state[l] = (h, c)
            inp = h
            step_trace.append((h, c))
        outs.append(inp)
        trace.append(step_trace)
    out = torch.stack(outs, dim=1)
    if return_all:
        return out, state, trace
    return out, state


def stack_backward(x_seq: Tensor, weights: Sequence[Tensor],
                   biases: Sequence[Optional[Tensor]], d_out: Tensor):
    """Explicit BPTT through :func:`stack_forward` from zero state.

    ``d_out`` is the gradient w.r.t. the top layer's ``h`` at every step
    ``[B,T,Ch_L,H,W]``.  Returns ``(dx_seq, [dW_l], [db_l])``.
    Layer wiring: ``dx`` of layer ``l`` at step ``t`` adds into ``dh`` of layer
    ``l-1`` at the same ``t`` (last paragraph).
    """
    n_layers = len(weights)
    b, t_steps, _, hh, ww = x_seq.shape
    # forward, keeping (input, h_prev, c_prev) per layer-step
    saved = []
    state = []
    for w in weights:
        ch = w.shape[0] // 4
        z = torch.zeros(b, ch, hh, ww, dtype=w.dtype)
        state.append((z, z.clone()))
    for t in range(t_steps):
        inp = x_seq[:, t]
        row = []
        for l in range(n_layers):
            h, c = state[l]
            row.append((inp, h, c))
            h, c = cell_forward(inp, h, c, weights[l], biases[l])
            state[l] = (h, c)
            inp = h
        saved.append(row)
    dws = [torch.zeros_like(w) for w in weights]
    dbs = [torch.zeros(w.shape[0], dtype=w.dtype) for w in weights]
    dh_carry = [torch.zeros_like(s[0]) for s in state]
    dc_carry = [torch.zeros_like(s[1]) for s in state]
    dx_seq = torch.zeros_like(x_seq)
    for t in reversed(range(t_steps)):
        d_from_above = d_out[:, t]
        for l in reversed(range(n_layers)):
            inp, h_prev, c_prev = saved[t][l]
            dh = dh_carry[l] + d_from_above
            r = cell_backward(inp, h_prev, c_prev, weights[l], biases[l], dh, dc_carry[l])
            dws[l] += r["dW"]
            dbs[l] += r["db"]
            dh_carry[l] = r["dh_prev"]
            dc_carry[l] = r["dc_prev"]
            d_from_above = r["dx"]
        dx_seq[:, t] = d_from_above
    return dx_seq, dws, dbs


def encoder_forecaster_forward(x_seq: Tensor, enc_w, enc_b, fc_w, fc_b, t_out: int):
    """Encoder-forecaster rollout (north_star extension; NO reference counterpart).

    Spec defined by this repo (parity vs this eager restatement only):
    the encoder stack consumes ``x_seq`` from zero state; the forecaster stack
    (own weights) starts from the encoder's final ``(h, c)`` per layer and runs
    ``t_out`` steps; its first layer has no input (``Cin = 0``, weight
    ``[4Ch, Ch, k, k]``), layer ``l > 0`` consumes ``h`` of layer ``l-1``.
    Returns the top layer's ``h`` for the ``t_out`` forecast steps.
    """
    _, state = stack_forward(x_seq, enc_w, enc_b)
    out, state = stack_forward(None, fc_w, fc_b, state=state, steps=t_out)
    return out, state


def add_coord_channels(x: Tensor) -> Tensor:
    """coordconv.py:3-10: append row / col ``linspace(0, 1)`` channels."""
    b, _, hh, ww = x.shape
    row = torch.linspace(0, 1, hh, dtype=x.dtype).view(1, 1, hh, 1).repeat(b, 1, 1, ww)
    col = torch.linspace(0, 1, ww, dtype=x.dtype).view(1, 1, 1, ww).repeat(b, 1, hh, 1)
    return torch.cat([x, row, col], dim=1)


def frontend_forward(frame: Tensor, w_init: Tensor, b_init: Optional[Tensor]) -> Tensor:
    """generator.py:166-168: ``x_t = relu(init_conv(add_coord_channels(frame_t)))``."""
    return F.relu(F.conv2d(add_coord_channels(frame), w_init, b_init, padding=1))


def nowcast_forward(frames: Tensor, w_init, b_init, enc_w, enc_b, fc_w, fc_b, w_head, b_head, t_out: int) -> Tensor:
    """Encoder-forecaster generator (repo-defined spec, see :func:`encoder_forecaster_forward`):
    frames [B,T_in,Cf,H,W] -> front-end per step -> encoder -> forecaster -> 1x1 head -> [B,T_out,1,H,W]."""
    b, t_in = frames.shape[:2]
    feats = torch.stack([frontend_forward(frames[:, t], w_init, b_init) for t in range(t_in)], dim=1)
    h_top, _ = encoder_forecaster_forward(feats, enc_w, enc_b, fc_w, fc_b, t_out)
    outs = [F.conv2d(h_top[:, t], w_head, b_head) for t in range(t_out)]
    return torch.stack(outs, dim=1)
