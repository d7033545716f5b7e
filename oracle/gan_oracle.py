"""Eager CPU spec of the clip discriminator and the GAN losses (TEST INFRASTRUCTURE ONLY).

**Parity unpinned against the reference by construction**: the reference has NO discriminator, no ``Conv3d`` and no
adversarial loss anywhere ("GAN" appears in its repo name and in ``main.py:13`` only -- SURVEY.md section 0).
``BASELINE.json:north_star`` nevertheless asks for "the discriminator's strided 2D/3D convolutions" on the same
implicit-GEMM core and for "generator/discriminator losses over a full T-step rollout", so this repo defines both, and
this file -- plain ``F.conv2d`` / ``F.conv3d`` / ``F.leaky_relu`` / ``F.binary_cross_entropy_with_logits`` -- is the
spec the CUDA path (``pl-convlstm-gan_b200/gan.py``) is checked against.  Every result checked against it is labelled
"parity vs repo spec", never "vs reference".

Discriminator (a 20-frame clip = conditioning frames ++ real or predicted frames, [N, T, 1, H, W]):
    conv1  Conv2d(1, 32, 3, stride 2, pad 1) on every frame          + LeakyReLU(0.2)   -> [N, T,   H/2, W/2, 32]
    conv2  Conv3d(32, 64, 3, stride (1, 2, 2), pad 1)                 + LeakyReLU(0.2)   -> [N, T,   H/4, W/4, 64]
    conv3  Conv3d(64, 128, 3, stride (2, 2, 2), pad 1)                + LeakyReLU(0.2)   -> [N, T/2, H/8, W/8, 128]
    score  Conv2d(128, 1, 3, pad 1) on every remaining frame; logit = mean over (T/2, H/8, W/8)
Losses (non-saturating GAN, as pix2pix-style conditional GANs use):
    L_D = BCE(D(real), 1) + BCE(D(fake), 0)            L_G = L1(fake, real) + lambda_adv * BCE(D(fake), 1)
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
WIDTHS = (32, 64, 128)
SLOPE = 0.2


def make_discriminator_params(seed: int = 0, in_channels: int = 1, widths=WIDTHS, dtype=torch.float32) -> Dict[str, Tensor]:
    """torch's default Conv init (kaiming-uniform a=sqrt(5) == U(+-1/sqrt(fan_in)) for weight and bias)."""
    g = torch.Generator().manual_seed(seed)

    def mk(shape):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = 1.0 / fan_in ** 0.5
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype), \
               ((torch.rand(shape[0], generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    w0, w1, w2 = widths
    p = {}
    p["conv1.weight"], p["conv1.bias"] = mk((w0, in_channels, 3, 3))
    p["conv2.weight"], p["conv2.bias"] = mk((w1, w0, 3, 3, 3))
    p["conv3.weight"], p["conv3.bias"] = mk((w2, w1, 3, 3, 3))
    p["score.weight"], p["score.bias"] = mk((1, w2, 3, 3))
    return p


def _lrelu(x: Tensor, mask):
    """LeakyReLU; with ``mask`` (bool, same shape) the positive set is prescribed instead of taken from sign(x) -- used by
    gradient checks to condition the spec on the activation pattern of the implementation under test (at an activation
    within rounding error of 0 the kink makes the gradient ill-defined, and bf16 vs fp64 may land on opposite sides)."""
    if mask is None:
        return F.leaky_relu(x, SLOPE)
    return torch.where(mask, x, SLOPE * x)


def discriminator_features(clips: Tensor, p: Dict[str, Tensor], masks=None):
    """Every layer's activation, in torch layout: conv1 [N*T,32,h,w]; conv2 / conv3 [N,C,T,h,w]; score [N*T',1,h,w].
    masks: optional (m1, m2, m3) bool tensors in the same layouts prescribing the LeakyReLU positive sets."""
    n, t, c, hh, ww = clips.shape
    m1, m2, m3 = masks if masks is not None else (None, None, None)
    a1 = _lrelu(F.conv2d(clips.reshape(n * t, c, hh, ww), p["conv1.weight"], p["conv1.bias"], stride=2, padding=1), m1)
    x = a1.view(n, t, *a1.shape[1:]).permute(0, 2, 1, 3, 4)                       # [N, C, T, h, w]
    a2 = _lrelu(F.conv3d(x, p["conv2.weight"], p["conv2.bias"], stride=(1, 2, 2), padding=1), m2)
    a3 = _lrelu(F.conv3d(a2, p["conv3.weight"], p["conv3.bias"], stride=(2, 2, 2), padding=1), m3)
    t2 = a3.shape[2]
    y = a3.permute(0, 2, 1, 3, 4).reshape(n * t2, a3.shape[1], *a3.shape[3:])
    s = F.conv2d(y, p["score.weight"], p["score.bias"], padding=1)
    return a1, a2, a3, s


def discriminator_forward(clips: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """clips [N, T, 1, H, W] -> logits [N]."""
    s = discriminator_features(clips, p)[3]
    return s.reshape(clips.shape[0], -1).mean(1)


def d_loss(logits: Tensor, n_real: int) -> Tensor:
    """logits = [real samples; fake samples]."""
    real, fake = logits[:n_real], logits[n_real:]
    return F.binary_cross_entropy_with_logits(real, torch.ones_like(real)) + \
        F.binary_cross_entropy_with_logits(fake, torch.zeros_like(fake))


def g_adv_loss(logits_fake: Tensor) -> Tensor:
    return F.binary_cross_entropy_with_logits(logits_fake, torch.ones_like(logits_fake))
