"""CPU restatement of the reference's CombinedLoss (TEST INFRASTRUCTURE ONLY; SURVEY.md section 8f "next-2").

Follows ``src/losses/combined_loss.py``:
  * sample weights            :22-59   (``log``: 1 + log1p(rain); ``sqrt``; ``stratified``; off)
  * conservation (area pool)  :64-74
  * station point supervision :79-141  (coords scaled with cell-centre alignment, NaN observations dropped)
  * spatial gradient          :146-155
  * temporal consistency      :160-168
  * weighted total            :173-191 (defaults 1.0, 1.0, 0.1, 0.05)
Pinned by tests/golden/generator_*.npz (loss terms computed by the unmodified reference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def sample_weights(rain, strategy="log", use_weighted=True):
    if not use_weighted:
        return torch.ones_like(rain)
    if strategy == "log":
        return 1.0 + torch.log1p(rain)                                   # combined_loss.py:42
    if strategy == "sqrt":
        return 1.0 + torch.sqrt(rain)                                    # combined_loss.py:54
    if strategy == "stratified":                                         # combined_loss.py:46-49
        w = torch.ones_like(rain)
        for thr, val in ((10, 2.0), (25, 3.0), (50, 5.0)):
            w = torch.where(rain >= thr, torch.full_like(rain, val), w)
        return w
    return torch.ones_like(rain)


def conservation(pred, lr_input):
    b, t, c, h, w = pred.shape
    hl, wl = lr_input.shape[-2:]
    pooled = F.interpolate(pred.reshape(b * t, c, h, w), size=(hl, wl), mode="area").reshape(b, t, c, hl, wl)
    return (pooled - lr_input).abs().mean()                              # nn.L1Loss, combined_loss.py:74


def point_supervision(pred, s_coords, s_values, scale_factor=1.0, strategy="log", use_weighted=True):
    if s_values is None or s_coords.numel() == 0:
        return pred.new_zeros(())
    b, t, _, h, w = pred.shape
    coords = s_coords[0] if s_coords.dim() == 3 else s_coords
    scaled = ((coords.float() + 0.5) * scale_factor - 0.5).long()        # combined_loss.py:97
    rows, cols = scaled[:, 0], scaled[:, 1]
    ok = (rows >= 0) & (rows < h) & (cols >= 0) & (cols < w)
    if int(ok.sum()) == 0:
        return pred.new_zeros(())
    at_st = pred[:, :, 0][:, :, rows[ok], cols[ok]]                      # [B, T, n_valid]
    obs = s_values[:, :, ok] if s_values.dim() == 3 else s_values[:, ok].unsqueeze(0).expand(b, -1, -1)
    m = ~torch.isnan(obs)
    if int(m.sum()) == 0:
        return pred.new_zeros(())
    wts = sample_weights(obs[m], strategy, use_weighted)
    return ((at_st[m] - obs[m]).abs() * wts).mean()                      # combined_loss.py:134-139


def spatial_gradient(pred):
    gx = (pred[..., :, :-1] - pred[..., :, 1:]).abs().mean()
    gy = (pred[..., :-1, :] - pred[..., 1:, :]).abs().mean()
    return gx + gy                                                       # combined_loss.py:150-155


def temporal_consistency(pred):
    return (pred[:, :-1] - pred[:, 1:]).abs().mean()                     # combined_loss.py:166-168


def combined_loss(pred, lr_input, s_coords, s_values, scale_factor=1.0, lambdas=(1.0, 1.0, 0.1, 0.05),
                  strategy="log", use_weighted=True):
    parts = {
        "point": point_supervision(pred, s_coords, s_values, scale_factor, strategy, use_weighted),
        "conserve": conservation(pred, lr_input),
        "smooth": spatial_gradient(pred),
        "temporal": temporal_consistency(pred),
    }
    lp, lc, ls, lt = lambdas
    total = lp * parts["point"] + lc * parts["conserve"] + ls * parts["smooth"] + lt * parts["temporal"]
    return total, parts                                                  # combined_loss.py:179-191
