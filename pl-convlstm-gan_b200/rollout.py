"""Preallocated inference rollout of the encoder-forecaster generator (BASELINE.json configs[1]).

frames [B,T_in,Cf,H,W] fp32 -> front-end conv (generator.py:166-168) -> encoder ConvLSTM stack over T_in
steps (generator.py:156-171 wiring) -> forecaster stack over T_out steps (input-less first layer) -> 1x1 head ->
frames [T_out,B,H,W] fp32.  All state lives in buffers allocated once; one cell step = one kernel launch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import functional as F
from .nn import ConvLSTMStack, _MODES

Tensor = torch.Tensor


class NowcastGenerator(nn.Module):
    """Encoder-forecaster generator (north_star extension; parity vs the repo's own eager spec in the tests,
    not vs the reference, which has no such model).

    Parameter names follow the reference where a counterpart exists: ``init_conv`` (generator.py:50-55),
    ``encoder.cells[l].conv`` / ``forecaster.cells[l].conv`` (ConvLSTMCell.conv)."""

    def __init__(self, in_channels: int = 1, hidden_dims: Sequence[int] = (64, 64), kernel_size: int = 3,
                 t_in: int = 10, t_out: int = 10, mode: str = "bf16"):
        super().__init__()
        self.in_channels, self.hidden_dims = in_channels, list(hidden_dims)
        self.t_in, self.t_out, self.mode = t_in, t_out, mode
        hd0 = self.hidden_dims[0]
        self.init_conv = nn.Conv2d(in_channels + 2, hd0, 3, padding=1)      # parameter holder
        self.encoder = ConvLSTMStack(hd0, self.hidden_dims, kernel_size, True, mode)
        self.forecaster = ConvLSTMStack(0, self.hidden_dims, kernel_size, True, mode)
        self.head = nn.Conv2d(self.hidden_dims[-1], 1, 1)                   # parameter holder

    def forward(self, frames: Tensor) -> Tensor:
        """Differentiable rollout (training): frames [B,T_in,Cf,H,W] fp32 -> predicted frames [B,T_out,1,H,W] fp32.
        bf16 mode only (every layer runs in libplc.so: front-end on the tensor-core conv, fused BPTT, 1x1 head)."""
        if self.mode != "bf16":
            raise RuntimeError("NowcastGenerator.forward is implemented for mode='bf16'")
        B, T, Cf, H, W = frames.shape
        cp0 = self._init_cp()
        if (cp0.Cout == 64 and cp0.k == 3 and cp0.cin_p == 8 and cp0.Cin * 9 <= 32 and not frames.requires_grad
                and F.frontend_tc_supported(Cf, 64, 64)):
            # fused front-end: coordinate planes + im2col inside the kernel, straight from the fp32 frames
            feat = F.frontend_tc_train(frames.contiguous(), cp0).view(T, B, H, W, 64)   # generator.py:166-168
        else:
            x = F.frames_to_nhwc(frames.contiguous(), cp0.cin_p)                   # + coord planes (coordconv.py:3-10)
            feat = F.conv2d_same(x, cp0).view(T, B, H, W, cp0.cout_p)               # generator.py:166-168
        _, state = self.encoder.run_seq(feat)
        out, _ = self.forecaster.run_seq(None, state, steps=self.t_out)            # [T_out,B,H,W,Ch]
        y = F.head(out, self.head.weight, self.head.bias, _MODES[self.mode])       # [T_out,B,H,W] fp32
        return y.permute(1, 0, 2, 3).unsqueeze(2)

    def _init_cp(self):
        cp = getattr(self, "_cp_init", None)
        if cp is None or cp.conv is not self.init_conv:
            cp = F.ConvParams(self.init_conv, relu=True)
            object.__setattr__(self, "_cp_init", cp)
        return cp

    def cell_flops_per_sequence(self, H: int, W: int) -> float:
        """Algorithmic FLOPs (2*M*N*K of every gate conv) for ONE sequence through encoder + forecaster."""
        tot = 0.0
        for stack, steps in ((self.encoder, self.t_in), (self.forecaster, self.t_out)):
            for cell in stack.cells:
                k = cell.kernel_size
                tot += steps * 2.0 * H * W * (cell.input_dim + cell.hidden_dim) * k * k * 4 * cell.hidden_dim
        return tot


class NowcastRunner:
    """Inference engine for :class:`NowcastGenerator` with every buffer preallocated.

    The kernel-layout weight images (cells, front-end, head) are snapshots of the model's parameters; every ``run``
    compares the parameters' version counters and the packed-weight generation (advanced by optimizer steps) with the
    snapshot's and re-packs when they differ, so a runner survives training steps and ``load_state_dict``.  A captured
    CUDA graph (:meth:`capture`) replays against the snapshot it was captured with: re-capture after weight updates."""

    def __init__(self, model: NowcastGenerator, B: int, H: int, W: int, device):
        self.m, self.B, self.H, self.W, self.dev = model, B, H, W, device
        self.mode = _MODES[model.mode]
        adt = torch.bfloat16 if model.mode == "bf16" else torch.float32
        hd = model.hidden_dims
        L = len(hd)
        self.feat = torch.zeros(model.t_in * B, H, W, model.encoder.cells[0].working_cin, dtype=adt, device=device)
        # bf16 mode: the front-end conv runs on the tensor-core conv (init_conv + ReLU), fed by a coord-plane prep kernel
        self.tc_frontend = model.mode == "bf16" and model.hidden_dims[0] % 8 == 0
        # ... or, for <= 3 frame channels into 64 features, on the dedicated kernel with in-kernel im2col
        self.fused_frontend = (self.tc_frontend and
                               F.frontend_tc_supported(model.in_channels, hd[0], self.feat.shape[-1]))
        if self.fused_frontend:
            self._fe_w = model.init_conv.weight.detach().to(torch.float32).contiguous()
            self._fe_b = None if model.init_conv.bias is None else model.init_conv.bias.detach().float().contiguous()
        elif self.tc_frontend:
            self.cp0 = model._init_cp()
            self.x8 = torch.zeros(model.t_in * B, H, W, self.cp0.cin_p, dtype=torch.bfloat16, device=device)
        # per layer: two h buffers (ping-pong; halo reads forbid in-place) and one c buffer (in-place is safe)
        self.h = [[torch.zeros(B, H, W, hd[l], dtype=adt, device=device) for _ in range(2)] for l in range(L)]
        self.c = [torch.zeros(B, H, W, hd[l], dtype=torch.float32, device=device) for l in range(L)]
        self.h_top = torch.zeros(model.t_out, B, H, W, hd[-1], dtype=adt, device=device)
        self.out = torch.zeros(model.t_out, B, H, W, dtype=torch.float32, device=device)
        self._weights_key = None
        self.refresh_weights()
        self.zero_state = [F.zero_state_supported(pw) for pw in self.enc_pw]
        self.cell_launches_per_run = L * (model.t_in + model.t_out)
        self.launches_per_run = (2 if (self.tc_frontend and not self.fused_frontend) else 1) + \
            self.cell_launches_per_run + 1

    def _key(self):
        from . import _lib
        return (_lib.weight_generation(), tuple(p._version for p in self.m.parameters()),
                tuple(p.data_ptr() for p in self.m.parameters()))

    def refresh_weights(self) -> bool:
        """Re-snapshot the packed weight images if the model's parameters changed since the last snapshot."""
        key = self._key()
        if key == self._weights_key:
            return False
        m = self.m
        if getattr(self, "fused_frontend", False):
            self._fe_w = m.init_conv.weight.detach().to(torch.float32).contiguous()
            self._fe_b = None if m.init_conv.bias is None else m.init_conv.bias.detach().float().contiguous()
        self.enc_pw = [c._packed(False) for c in m.encoder.cells]
        self.fc_pw = [c._packed(False) for c in m.forecaster.cells]
        self._weights_key = key
        return True

    # ------------------------------------------------------------------ CUDA graph (launch-bound shapes)
    def capture(self, frames_like: Tensor):
        """Capture one whole rollout (front-end, T x L cell steps, head: 2 + L*(T_in+T_out) launches + state
        resets) into a CUDA graph.  Worth it when a cell step is shorter than its host-side launch cost (small
        frames / narrow hidden sizes); the tensor maps are kernel parameters, so they are baked into the graph and
        every buffer involved is one of the runner's preallocated tensors.  Returns self; use :meth:`replay`."""
        self._g_in = torch.empty_like(frames_like)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                         # warm-up outside capture (lazy module loading etc.)
            self.run(self._g_in)
        torch.cuda.current_stream().wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_out = self.run(self._g_in)
        return self

    def replay(self, frames: Tensor) -> Tensor:
        """Run the captured rollout on new frames (copied into the graph's static input buffer)."""
        self._g_in.copy_(frames, non_blocking=True)
        self._graph.replay()
        return self._g_out

    def _cell(self, pw, x, h_prev, c, h_out, events):
        if events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if h_prev is None:
            F.cell_forward_zero_state(x, pw, h_out=h_out, c_out=c)
        else:
            F.cell_forward(x, h_prev, c, pw, h_out=h_out, c_out=c)
        if events is not None:
            e1.record()
            events.append((e0, e1, pw))

    @torch.no_grad()
    def run(self, frames: Tensor, events: Optional[List] = None, out: Optional[Tensor] = None) -> Tensor:
        """frames [B,T_in,Cf,H,W] fp32 on device.  Returns predicted frames [T_out,B,H,W] fp32 (``out`` or the
        runner's own device buffer).  ``events``: optional list collecting (start_event, end_event, packed_weights)
        per cell launch."""
        m, B, L = self.m, self.B, len(self.c)
        T_in, T_out = m.t_in, m.t_out
        if not torch.cuda.is_current_stream_capturing():
            self.refresh_weights()
        # front-end for all T_in steps at once (T-major batch [T*B, ...])
        if self.fused_frontend:
            F.frontend_tc(frames, self._fe_w, self._fe_b, self.feat)
        elif self.tc_frontend:
            F.frames_to_nhwc(frames, self.cp0.cin_p, out=self.x8)
            F.conv2d_same_into(self.x8, self.cp0, self.feat)
        else:
            fr = frames.transpose(0, 1).reshape(T_in * B, m.in_channels, self.H, self.W).contiguous()
            F.frontend_forward(fr, m.init_conv.weight, m.init_conv.bias, self.mode, c_stride=self.feat.shape[-1],
                               out=self.feat)
        # generator.py:156-160: zero initial state.  Where the library offers the zero-state form, the first step skips
        # the h taps and the c read instead of zero-filling the state buffers and multiplying by them.
        for l in range(L):
            if not self.zero_state[l]:
                self.h[l][0].zero_()
                self.c[l].zero_()
        cur = [0] * L
        for t in range(T_in):                                # generator.py:164
            x = self.feat[t * B:(t + 1) * B]
            for l in range(L):                               # generator.py:170-171
                dst = self.h[l][cur[l] ^ 1]
                self._cell(self.enc_pw[l], x, None if (t == 0 and self.zero_state[l]) else self.h[l][cur[l]],
                           self.c[l], dst, events)
                cur[l] ^= 1
                x = dst
        for t in range(T_out):
            x = None                                         # forecaster layer 0 has no input tensor
            for l in range(L):
                if l == L - 1:                               # top layer writes straight into the output ring
                    src = self.h[l][cur[l]] if t == 0 else self.h_top[t - 1]
                    dst = self.h_top[t]
                    self._cell(self.fc_pw[l], x, src, self.c[l], dst, events)
                else:
                    dst = self.h[l][cur[l] ^ 1]
                    self._cell(self.fc_pw[l], x, self.h[l][cur[l]], self.c[l], dst, events)
                    cur[l] ^= 1
                x = dst
        out = self.out if out is None else out
        F.head_forward(self.h_top, m.head.weight, m.head.bias, self.mode, out=out)
        return out
