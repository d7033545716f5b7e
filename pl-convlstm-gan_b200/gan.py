"""Clip discriminator and the GAN training step (BASELINE.json configs[2]: "full generator+discriminator GAN training").

No reference counterpart: the reference has no discriminator, no Conv3d and no adversarial loss (SURVEY.md section 0;
"GAN" is in its name and in main.py:13 only).  The north_star still names "the discriminator's strided 2D/3D
convolutions [that] reuse the same implicit-GEMM core", so this repo defines the model -- eager spec:
``oracle/gan_oracle.py``; every parity statement is "vs repo spec".  All four convolutions run in libplc.so
(``plc_convnd_fwd`` / ``plc_convnd_bwd``: conv_igemm_tc_kernel with strided and 5-D tensor maps, bias + LeakyReLU in the
epilogue); the step order of each optimizer follows the reference loop (src/training/trainer.py:290-315):
zero_grad -> forward -> loss -> NaN-skip -> backward -> (all-reduce) -> clip 0.5 -> Adam.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F
from .training import TrainStep

Tensor = torch.Tensor


class Discriminator(nn.Module):
    """clip [N, T, C, H, W] fp32 (conditioning frames ++ real or predicted frames) -> logits [N].

    conv1 Conv2d(C,32,3,s2) per frame -> conv2 Conv3d(32,64,3,s(1,2,2)) -> conv3 Conv3d(64,128,3,s(2,2,2)), LeakyReLU(0.2)
    after each -> score Conv2d(128,1,3) per frame -> mean.  The nn.Conv2d / nn.Conv3d modules are PARAMETER HOLDERS
    (state_dict keys conv1/conv2/conv3/score .weight/.bias, torch default init); they are never called.
    conv1 runs in plc_frameconv_* (reads the fp32 frames directly) when its shape allows, else on the implicit-GEMM
    core like the other three."""

    def __init__(self, in_channels: int = 1, widths: Sequence[int] = (32, 64, 128), slope: float = 0.2):
        super().__init__()
        w0, w1, w2 = widths
        self.in_channels, self.slope = in_channels, slope
        self.conv1 = nn.Conv2d(in_channels, w0, 3, stride=2, padding=1)
        self.conv2 = nn.Conv3d(w0, w1, 3, stride=(1, 2, 2), padding=1)
        self.conv3 = nn.Conv3d(w1, w2, 3, stride=(2, 2, 2), padding=1)
        self.score = nn.Conv2d(w2, 1, 3, padding=1)
        self._cps = None

    def _params(self):
        cps = self._cps
        if cps is None or cps[0].conv is not self.conv1:
            cps = self._cps = (F.ConvNdParams(self.conv1, act=2, slope=self.slope),
                               F.ConvNdParams(self.conv2, act=2, slope=self.slope),
                               F.ConvNdParams(self.conv3, act=2, slope=self.slope),
                               F.ConvNdParams(self.score, act=0))
        return cps

    def features(self, clip: Tensor):
        """Working-layout activations of every layer: [N*T,1,h,w,32], [N,T,h,w,64], [N,T/2,h,w,128], [N*T/2,1,h,w,8]."""
        if not clip.is_cuda:
            raise RuntimeError("Discriminator (plconv) has no CPU path")
        n, t, c, hh, ww = clip.shape
        cp1, cp2, cp3, cp4 = self._params()
        if F.frameconv_supported(self.conv1):
            # first layer straight from the fp32 frames (plc_frameconv_*: HBM-bound SIMT kernels, no pad / cast pass)
            a1 = F.frameconv(clip.reshape(n * t, c, hh, ww).float().contiguous(), self.conv1, act=2, slope=self.slope)
            a1 = a1.unsqueeze(1)
        else:
            x = clip.permute(0, 1, 3, 4, 2)                               # [N,T,H,W,C] (a view when C == 1)
            x = TF.pad(x, (0, cp1.cin_p - c)).to(torch.bfloat16).contiguous()
            a1 = F.convnd(x.view(n * t, 1, hh, ww, cp1.cin_p), cp1)       # strided 2-D conv, every frame
        a2 = F.convnd(a1.view(n, t, *a1.shape[2:]), cp2)                  # strided 3-D conv (time stride 1)
        a3 = F.convnd(a2, cp3)                                            # strided 3-D conv (time stride 2)
        s = F.convnd(a3.view(n * a3.shape[1], 1, *a3.shape[2:]), cp4)     # 3x3 score conv, every remaining frame
        return a1, a2, a3, s

    def forward(self, clip: Tensor) -> Tensor:
        s = self.features(clip)[3]
        return s[..., 0].float().reshape(clip.shape[0], -1).mean(1)


def bce_with_logits(logits: Tensor, target: float) -> Tensor:
    """mean BCE against a constant label: softplus(-x) for 1, softplus(x) for 0."""
    return TF.softplus(-logits).mean() if target == 1.0 else TF.softplus(logits).mean()


class GanTrainStep:
    """One GAN training step = a discriminator step followed by a generator step.

        fake = G(frames)
        D:  L_D = BCE(D(frames ++ target), 1) + BCE(D(frames ++ fake.detach()), 0)   (one batched D forward)
        G:  L_G = L1(fake, target) + lambda_adv * BCE(D(frames ++ fake), 1)          (through the updated D; D's
                                                                                      parameters take no gradient)
    Each half is a :class:`plconv.training.TrainStep`: bucketed NCCL all-reduce overlapped with the backward pass,
    clip 0.5, fused Adam, device-side NaN-skip -- no host synchronisation anywhere in the step."""

    def __init__(self, gen: nn.Module, disc: Discriminator, lr_g: float = 5e-4, lr_d: float = 2e-4,
                 lambda_adv: float = 0.05, grad_clip_norm: float = 0.5, process_group=None):
        self.gen, self.disc, self.lambda_adv = gen, disc, lambda_adv
        if hasattr(gen, "forecaster") and hasattr(gen, "encoder"):
            # buckets in the order BPTT finishes them: forecaster cells, encoder cells, then front-end + head
            claimed = set()
            groups = []
            for stack in (gen.forecaster, gen.encoder):
                for cell in reversed(list(stack.cells)):
                    groups.append(list(cell.parameters()))
                    claimed.update(id(p) for p in groups[-1])
            rest = [p for p in gen.parameters() if id(p) not in claimed]
            if rest:
                groups.append(rest)
        else:
            groups = [list(gen.parameters())]
        self.g = TrainStep(gen, groups, lr=lr_g, grad_clip_norm=grad_clip_norm, process_group=process_group)
        self.d = TrainStep(disc, [list(disc.parameters())], lr=lr_d, grad_clip_norm=grad_clip_norm,
                           process_group=process_group, betas=(0.5, 0.999))
        self.last = {}

    def losses(self, frames: Tensor, target: Tensor, fake: Tensor):
        """(L_D, L1, adversarial term of G) for given tensors -- used by the parity tests."""
        b = frames.shape[0]
        clips = torch.cat([torch.cat([frames, target], 1), torch.cat([frames, fake.detach()], 1)], 0)
        logits = self.disc(clips)
        d_loss = bce_with_logits(logits[:b], 1.0) + bce_with_logits(logits[b:], 0.0)
        adv = bce_with_logits(self.disc(torch.cat([frames, fake], 1)), 1.0)
        return d_loss, (fake - target).abs().mean(), adv

    def __call__(self, frames: Tensor, target: Tensor) -> Tensor:
        b = frames.shape[0]
        fake = self.gen(frames)                                            # graph kept for the G step
        # ---- discriminator step
        self.d.zero_grad()
        clips = torch.cat([torch.cat([frames, target], 1), torch.cat([frames, fake.detach()], 1)], 0)
        logits = self.disc(clips)
        d_loss = bce_with_logits(logits[:b], 1.0) + bce_with_logits(logits[b:], 0.0)
        self.d.backward_and_step(d_loss)
        # ---- generator step (adversarial term through the updated D, whose parameters take no gradient here)
        self.g.zero_grad()
        d_params = list(self.disc.parameters())
        for p in d_params:
            p.requires_grad_(False)
        adv = bce_with_logits(self.disc(torch.cat([frames, fake], 1)), 1.0)
        for p in d_params:
            p.requires_grad_(True)
        l1 = (fake - target).abs().mean()
        g_loss = l1 + self.lambda_adv * adv
        self.g.backward_and_step(g_loss)
        self.last = {"d_loss": d_loss.detach(), "l1": l1.detach(), "adv": adv.detach()}
        return g_loss.detach()
