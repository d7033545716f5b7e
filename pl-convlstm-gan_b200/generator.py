"""Drop-in for the reference ``Generator`` (src/models/generator.py:31-205) built on the native ConvLSTM cells.

Same constructor, same ``forward(rain_lr, dem, lu, input_grid_size=None)``, same parameter names, so a reference
checkpoint (``best_model.pth['model_state_dict']``, trainer.py:410-417) loads unchanged:
    init_conv.*, cell1.conv.*, cell2.conv.*, dem_attn.conv.{0,2}.*, lu_attn.conv.{0,2}.*,
    upsample_blocks.{i}.conv.*, post_process.{0,2}.*
The recurrence (cell1/cell2 over T steps, generator.py:156-171) runs in libplc.so.  The NON-recurrent body
(front-end conv, PixelShuffle upsampling, DEM/LU gating, output head) is SURVEY.md section 8f "next-1": here it is plain
PyTorch in fp32 mode and NATIVE in bf16 mode: init_conv(+ReLU), the UpsampleBlock convs with PixelShuffle(2)+ReLU
fused into the store, and both post_process convs run on the same tcgen05 implicit-GEMM core (plc_conv_fwd/bwd,
NHWC bf16), batched over T (no recurrence there) with the time-invariant attention gates hoisted out of the T loop
(attention.py:13,26 depend only on static inputs; they are computed once per forward with torch ops).
Like the reference, ``upsample_blocks`` are created on first forward (generator.py:129-130; SURVEY.md section 5 gotcha 1);
call ``materialize(scale)`` before ``load_state_dict`` when loading a checkpoint into a fresh model.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as PF
from .nn import ConvLSTMCell, _StackRolloutFn


def _coord_channels(x: torch.Tensor) -> torch.Tensor:
    """coordconv.py:3-10: append row and column linspace(0, 1) planes."""
    b, _, h, w = x.shape
    rows = torch.linspace(0, 1, h, device=x.device, dtype=x.dtype).view(1, 1, h, 1).expand(b, 1, h, w)
    cols = torch.linspace(0, 1, w, device=x.device, dtype=x.dtype).view(1, 1, 1, w).expand(b, 1, h, w)
    return torch.cat([x, rows, cols], dim=1)


class _Gate(nn.Module):
    """DEMAttention / LUAttention (attention.py:3-26): x * sigmoid(conv1x1(relu(conv3x3(static)))).
    ``gate(static)`` returns the multiplicative map so callers can compute it once per forward."""

    def __init__(self, channels: int, static_channels: int):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(static_channels, channels // 2, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(channels // 2, channels, 1), nn.Sigmoid())

    def gate(self, static: torch.Tensor) -> torch.Tensor:
        return self.conv(static)

    def forward(self, x, static):
        return x * self.gate(static)


class _Up2(nn.Module):
    """UpsampleBlock (generator.py:10-28): conv3x3 C -> 4C, PixelShuffle(2), ReLU."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels * 4, 3, padding=1)

    def forward(self, x):
        return F.relu(F.pixel_shuffle(self.conv(x), 2))


class Generator(nn.Module):
    def __init__(self, in_channels: int = 1, dem_channels: int = 1, lu_channels: int = 0,
                 hidden_dims: Sequence[int] = (32, 64), target_grid_size=None, scale_factor=None, mode: str = "bf16"):
        super().__init__()
        self.hidden_dims = list(hidden_dims)
        self.lu_channels = lu_channels
        self.target_grid_size = target_grid_size if target_grid_size is not None else None
        self.scale_factor = None if target_grid_size is not None else scale_factor      # generator.py:38-48
        self.target_size = None
        hd0, hd1 = self.hidden_dims
        self.init_conv = nn.Conv2d(in_channels + 2, hd0, 3, padding=1)                 # generator.py:50-55
        self.cell1 = ConvLSTMCell(hd0, hd0, mode=mode)                                 # generator.py:57
        self.cell2 = ConvLSTMCell(hd0, hd1, mode=mode)                                 # generator.py:58
        self.dem_attn = _Gate(hd1, dem_channels)                                       # generator.py:60
        self.lu_attn = _Gate(hd1, lu_channels)                                         # generator.py:61
        self.upsample_blocks = None                                                    # generator.py:64 (lazy)
        self.post_process = nn.Sequential(nn.Conv2d(hd1, 32, 3, padding=1), nn.ReLU(inplace=True),
                                          nn.Conv2d(32, 1, 3, padding=1))              # generator.py:67-71
        self.mode = mode
        self._native = {}      # name -> functional.ConvParams (kernel-ready images of the plain convs, bf16 mode)

    def _cp(self, name: str, conv: nn.Conv2d, relu: bool, shuffle: bool = False, out_f32: bool = False):
        cp = self._native.get(name)
        if cp is None or cp.conv is not conv:
            cp = PF.ConvParams(conv, relu=relu, pixel_shuffle=shuffle, out_f32=out_f32)
            self._native[name] = cp
        return cp

    def materialize(self, scale_factor: int, device=None) -> float:
        """Create the x2 upsample blocks for an integer scale (generator.py:73-92); returns the residual factor."""
        blocks, f = [], int(scale_factor)
        while f >= 2:
            blocks.append(_Up2(self.hidden_dims[1]))
            f //= 2
        self.upsample_blocks = nn.ModuleList(blocks)
        if device is not None:
            self.upsample_blocks.to(device)
        return scale_factor / (2 ** len(blocks)) if scale_factor > 1 else 1

    def forward(self, rain_lr: torch.Tensor, dem: torch.Tensor, lu: torch.Tensor,
                input_grid_size: Optional[Tuple[float, float]] = None) -> torch.Tensor:
        B, T, C, H, W = rain_lr.shape
        dev = rain_lr.device
        # ---- target resolution (generator.py:106-126)
        if self.target_grid_size is not None and input_grid_size is not None:
            sw = input_grid_size[0] / self.target_grid_size[0]
            sh = input_grid_size[1] / self.target_grid_size[1]
            self.target_size = (int(H * sh), int(W * sw))
            scale = max(sh, sw)
        elif self.scale_factor is not None:
            scale, self.target_size = self.scale_factor, None
        else:
            scale, self.target_size = 1, None
        if self.upsample_blocks is None:
            remaining = self.materialize(int(scale), dev)                               # generator.py:129-130
        else:
            remaining = scale / (2 ** len(self.upsample_blocks))
        final_hw = self.target_size if self.target_size is not None else (int(H * scale), int(W * scale))
        # ---- static maps, once per forward (generator.py:143-153); gates are time-invariant
        dem_hr = F.interpolate(dem, size=final_hw, mode="bilinear", align_corners=False)
        lu_hr = F.interpolate(lu, size=final_hw, mode="nearest")
        gate = self.dem_attn.gate(dem_hr) * self.lu_attn.gate(lu_hr)                    # generator.py:198-199

        if self.mode == "bf16":
            return self._forward_native(rain_lr, gate, remaining)

        # ---- front-end for all T frames at once (generator.py:166-168)
        x = F.relu(self.init_conv(_coord_channels(rain_lr.reshape(B * T, C, H, W))))
        hd0, hd1 = self.hidden_dims
        c1, c2 = self.cell1, self.cell2
        xw = x.view(B, T, hd0, H, W).permute(1, 0, 3, 4, 2)                             # [T,B,H,W,C]
        if c1.working_cin != hd0:
            xw = F.pad(xw, (0, c1.working_cin - hd0))
        xw = xw.to(c1.act_dtype).contiguous()
        # ---- the recurrence (generator.py:156-171): zero state, cell1 then cell2 per step, in libplc.so;
        #      one fused autograd node for all T steps (explicit BPTT in its backward)
        zeros = [torch.zeros(B, H, W, hd0, device=dev, dtype=c1.act_dtype),
                 torch.zeros(B, H, W, hd0, device=dev, dtype=torch.float32),
                 torch.zeros(B, H, W, hd1, device=dev, dtype=c2.act_dtype),
                 torch.zeros(B, H, W, hd1, device=dev, dtype=torch.float32)]
        outs = _StackRolloutFn.apply([c1, c2], T, xw, *zeros, c1.conv.weight, c1.conv.bias, c2.conv.weight,
                                     c2.conv.bias)
        tops = outs[0]                                                                  # [T,B,H,W,hd1]
        feat = tops.permute(1, 0, 4, 2, 3).reshape(B * T, hd1, H, W).to(torch.float32)

        # ---- tail, batched over T (generator.py:173-203)
        for blk in self.upsample_blocks:
            feat = blk(feat)
        if remaining > 1:
            feat = F.interpolate(feat, scale_factor=remaining, mode="bilinear", align_corners=False)
        if self.target_size is not None:
            feat = F.interpolate(feat, size=self.target_size, mode="bilinear", align_corners=False)
        hh, ww = feat.shape[-2:]
        feat = (feat.view(B, T, hd1, hh, ww) * gate.unsqueeze(1)).view(B * T, hd1, hh, ww)
        out = self.post_process(feat)
        return out.view(B, T, 1, hh, ww)                                                # generator.py:205

    # ---- inference through a CUDA graph (launch-bound shapes: the shipped 15x12 / hidden [16,32] configuration
    #      spends its time in ~100 kernel launches of a few microseconds each)
    @torch.no_grad()
    def forward_graphed(self, rain_lr: torch.Tensor, dem: torch.Tensor, lu: torch.Tensor) -> torch.Tensor:
        """``forward`` for inference, captured once per (input shapes, weight state) into a CUDA graph and replayed.

        The tensor maps of libplc.so's kernels are kernel parameters, so they are baked into the graph; the graph is
        re-captured when the parameters change (optimizer step, ``load_state_dict``): the key carries the packed-weight
        generation and every parameter's version counter.  Returns a tensor owned by the graph (overwritten by the
        next call with the same key): ``.clone()`` it to keep it."""
        from . import _lib
        key = (tuple(rain_lr.shape), tuple(dem.shape), tuple(lu.shape), str(rain_lr.device), _lib.weight_generation(),
               tuple(p._version for p in self.parameters()))
        cache = self.__dict__.setdefault("_graphs", {})
        ent = cache.get(key)
        if ent is None:
            cache.clear()                                    # one live graph per module: old weight states are dead
            static = [t.clone() for t in (rain_lr, dem, lu)]
            side = torch.cuda.Stream(device=rain_lr.device)
            side.wait_stream(torch.cuda.current_stream(rain_lr.device))
            with torch.cuda.stream(side):                    # warm-up outside capture: lazy blocks, weight packing
                self.forward(*static)
            torch.cuda.current_stream(rain_lr.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(*static)
            # the graph holds raw pointers into the packed-weight images: keep them alive with it (a later backward /
            # re-pack replaces the modules' cache entries, which must not free what the graph replays against)
            keep = [cp._cache for cp in self._native.values()] + [self.cell1._pack_cache, self.cell2._pack_cache]
            ent = cache[key] = (graph, static, out, keep)
        graph, static, out, _ = ent
        for dst, src in zip(static, (rain_lr, dem, lu)):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        return out

    def _forward_native(self, rain_lr: torch.Tensor, gate: torch.Tensor, remaining: float) -> torch.Tensor:
        """bf16 mode: the whole per-step body in libplc.so (NHWC bf16, T-major batch [T*B, ...])."""
        B, T, C, H, W = rain_lr.shape
        dev = rain_lr.device
        hd0, hd1 = self.hidden_dims
        c1, c2 = self.cell1, self.cell2
        # coord channels (coordconv.py:3-10) appended in NHWC, channels zero-padded to the kernel granularity
        cp0 = self._cp("init", self.init_conv, relu=True)
        if rain_lr.dtype == torch.float32 and not rain_lr.requires_grad:
            x = PF.frames_to_nhwc(rain_lr.contiguous(), cp0.cin_p)      # one kernel: layout + coord planes + pad + cast
        else:
            x = rain_lr.permute(1, 0, 3, 4, 2).reshape(T * B, H, W, C)
            rows = torch.linspace(0, 1, H, device=dev, dtype=x.dtype).view(1, H, 1, 1).expand(T * B, H, W, 1)
            cols = torch.linspace(0, 1, W, device=dev, dtype=x.dtype).view(1, 1, W, 1).expand(T * B, H, W, 1)
            x = F.pad(torch.cat([x, rows, cols], dim=-1), (0, cp0.cin_p - (C + 2))).to(torch.bfloat16).contiguous()
        feat0 = PF.conv2d_same(x, cp0)                                                  # generator.py:166-168
        xw = feat0.view(T, B, H, W, cp0.cout_p)
        if cp0.cout_p != c1.working_cin:
            xw = xw[..., :hd0]
            xw = F.pad(xw, (0, c1.working_cin - hd0)).contiguous()
        zeros = [torch.zeros(B, H, W, hd0, device=dev, dtype=c1.act_dtype),
                 torch.zeros(B, H, W, hd0, device=dev, dtype=torch.float32),
                 torch.zeros(B, H, W, hd1, device=dev, dtype=c2.act_dtype),
                 torch.zeros(B, H, W, hd1, device=dev, dtype=torch.float32)]
        outs = _StackRolloutFn.apply([c1, c2], T, xw, *zeros, c1.conv.weight, c1.conv.bias, c2.conv.weight,
                                     c2.conv.bias)                                      # generator.py:156-171
        feat = outs[0].reshape(T * B, H, W, hd1)
        for i, blk in enumerate(self.upsample_blocks):                                  # generator.py:174-176
            feat = PF.conv2d_same(feat, self._cp(f"up{i}", blk.conv, relu=True, shuffle=True))
        if remaining > 1 or self.target_size is not None:                               # generator.py:179-195
            f32 = feat.permute(0, 3, 1, 2).float()
            if remaining > 1:
                f32 = F.interpolate(f32, scale_factor=remaining, mode="bilinear", align_corners=False)
            if self.target_size is not None:
                f32 = F.interpolate(f32, size=self.target_size, mode="bilinear", align_corners=False)
            feat = f32.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        hh, ww = feat.shape[1:3]
        g = gate.permute(0, 2, 3, 1).to(torch.bfloat16)                                 # [B,hh,ww,hd1], time-invariant
        feat = (feat.view(T, B, hh, ww, hd1) * g.unsqueeze(0)).reshape(T * B, hh, ww, hd1)   # generator.py:198-199
        cpa = self._cp("post0", self.post_process[0], relu=True)
        cpb = self._cp("post2", self.post_process[2], relu=False, out_f32=True)   # predicted rain stays fp32
        y = PF.conv2d_same(feat.contiguous(), cpa)                                      # generator.py:202
        if cpa.cout_p != cpb.cin_p:
            y = F.pad(y[..., :cpa.Cout], (0, cpb.cin_p - cpa.Cout)).contiguous()
        y = PF.conv2d_same(y, cpb)[..., 0]                                              # [T*B,hh,ww]
        return y.float().view(T, B, hh, ww).permute(1, 0, 2, 3).unsqueeze(2)            # generator.py:205
