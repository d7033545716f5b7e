"""Drop-in for the reference's ``CombinedLoss`` (``src/losses/combined_loss.py:6-191``) over ``plc_combined_loss``.

Same constructor arguments, same ``forward(pred, lr_input, s_coords, s_values, scale_factor)`` returning
``(total, {"point", "conserve", "smooth", "temporal"})``; the four reductions and d total / d pred come from three
fused passes in libplc.so (csrc/loss.cuh) instead of ~25 eager ops + their autograd graph.  CUDA only, fp32.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from .functional import _call, _ptr

_WEIGHT_MODES = {"log": 1, "sqrt": 2, "stratified": 3}


class _Problem:
    """Validated, device-resident arguments of one loss evaluation (shared by forward and backward)."""

    def __init__(self, pred, lr_input, coords, svals, coord_scale, lambdas, weight_mode):
        if not pred.is_cuda:
            raise RuntimeError("CombinedLoss: pred must be a CUDA tensor: this library has no CPU path")
        if pred.dim() != 5 or pred.shape[2] != 1:
            raise RuntimeError(f"CombinedLoss: pred must be [B,T,1,H,W] (got {tuple(pred.shape)})")
        B, T, _, Hs, Ws = pred.shape
        H, W = lr_input.shape[-2:]
        if lr_input.numel() != B * T * H * W:
            raise RuntimeError(f"CombinedLoss: lr_input {tuple(lr_input.shape)} does not match pred {tuple(pred.shape)}")
        if Hs % H or Ws % W or Hs // H != Ws // W:
            raise RuntimeError(f"CombinedLoss: only integer upsampling ratios are supported (pred {Hs}x{Ws}, "
                               f"lr {H}x{W})")
        self.pred = pred.detach().to(torch.float32).contiguous()
        dev = self.pred.device
        self.lr = lr_input.detach().to(device=dev, dtype=torch.float32).contiguous()
        n_st, has_batch = 0, 0
        if svals is not None and coords is not None and coords.numel() > 0:          # combined_loss.py:83
            coords = (coords[0] if coords.dim() == 3 else coords).to(device=dev, dtype=torch.int64).contiguous()
            svals = svals.detach().to(device=dev, dtype=torch.float32).contiguous()
            n_st, has_batch = coords.shape[0], int(svals.dim() == 3)
            want = (B, T, n_st) if has_batch else (T, n_st)
            if tuple(svals.shape) != want:
                raise RuntimeError(f"CombinedLoss: s_values {tuple(svals.shape)} does not match {want}")
        else:
            coords = svals = None
        self.coords, self.svals = coords, svals
        self.desc = _lib.PlcLossDesc(B, T, H, W, Hs // H, n_st, has_batch, weight_mode, float(coord_scale),
                                     *map(float, lambdas))

    def run(self, want_grad: bool, grad_scale=None):
        lib = _lib.load()
        dev = self.pred.device
        ws = torch.empty(lib.plc_loss_workspace_bytes(ctypes.byref(self.desc)), dtype=torch.uint8, device=dev)
        terms = torch.empty(5, dtype=torch.float32, device=dev)
        dpred = torch.empty_like(self.pred) if want_grad else None
        _call(self.pred, lib.plc_combined_loss, "plc_combined_loss", ctypes.byref(self.desc), _ptr(self.pred),
              _ptr(self.lr), _ptr(self.coords), _ptr(self.svals), _ptr(ws), _ptr(terms), _ptr(dpred), _ptr(grad_scale))
        return terms, dpred


class _CombinedLossFn(torch.autograd.Function):
    """forward: loss values only (pred read once).  backward: the same kernels again, now storing
    ``g_total * d total / d pred`` -- the upstream gradient is a device scalar, so nothing synchronises and no
    full-size gradient buffer lives between forward and backward."""

    @staticmethod
    def forward(ctx, pred, lr_input, coords, svals, coord_scale, lambdas, weight_mode):
        prob = _Problem(pred, lr_input, coords, svals, coord_scale, lambdas, weight_mode)
        terms, _ = prob.run(want_grad=False)
        ctx.prob = prob if pred.requires_grad else None
        ctx.pred_dtype = pred.dtype
        ctx.mark_non_differentiable(terms)
        return terms[0].clone(), terms

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        if ctx.prob is None:
            return (None,) * 7
        g = g_total.detach().to(device=ctx.prob.pred.device, dtype=torch.float32).reshape(1).contiguous()
        _, dpred = ctx.prob.run(want_grad=True, grad_scale=g)
        return dpred.to(ctx.pred_dtype), None, None, None, None, None, None


class CombinedLoss(nn.Module):
    """``CombinedLoss(lambda_point, lambda_conserve, lambda_smooth, lambda_temporal, use_weighted_loss,
    weight_strategy)`` -- combined_loss.py:7-18.  The returned parts are detached (the reference only logs them,
    trainer.py:321-330); gradients flow through ``total``."""

    def __init__(self, lambda_point=1.0, lambda_conserve=1.0, lambda_smooth=0.1, lambda_temporal=0.05,
                 use_weighted_loss=True, weight_strategy="log"):
        super().__init__()
        self.lambda_point = lambda_point
        self.lambda_conserve = lambda_conserve
        self.lambda_smooth = lambda_smooth
        self.lambda_temporal = lambda_temporal
        self.use_weighted_loss = use_weighted_loss
        self.weight_strategy = weight_strategy

    def _weight_mode(self) -> int:
        # unknown strategies fall back to unit weights, as combined_loss.py:56-57 does
        return _WEIGHT_MODES.get(self.weight_strategy, 0) if self.use_weighted_loss else 0

    def forward(self, pred, lr_input, s_coords, s_values, scale_factor=1.0):
        lambdas = (self.lambda_point, self.lambda_conserve, self.lambda_smooth, self.lambda_temporal)
        total, terms = _CombinedLossFn.apply(pred, lr_input, s_coords, s_values, scale_factor, lambdas,
                                             self._weight_mode())
        return total, {"point": terms[1], "conserve": terms[2], "smooth": terms[3], "temporal": terms[4]}
