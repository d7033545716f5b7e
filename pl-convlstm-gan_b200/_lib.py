"""ctypes binding of libplc.so (C ABI declared in include/plc.h).

There is deliberately NO fallback: if the library cannot be built/loaded, or a call returns a
non-zero status, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLC_LIB") or os.path.join(HERE, "libplc.so")   # PLC_LIB: developer override

PLC_MODE_BF16_TC = 0
PLC_MODE_FP32 = 1
PLC_PACK_FWD = 0
PLC_PACK_DGRAD = 1


class PlcCellDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "H", "W", "Cin", "Ch", "k", "mode", "has_bias")]


class PlcConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "H", "W", "Cin", "Cout", "k", "relu", "pixel_shuffle", "has_bias")]


class PlcLossDesc(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in ("B", "T", "H", "W", "scale", "n_stations", "svals_has_batch",
                                               "weight_mode")] +
                [(n, ctypes.c_float) for n in ("coord_scale", "lambda_point", "lambda_conserve", "lambda_smooth",
                                               "lambda_temporal")])


class PlcConvNdDesc(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in ("B", "T", "H", "W", "Cin", "Cout", "kt", "k", "stride_t", "stride", "act")] +
                [("slope", ctypes.c_float), ("has_bias", ctypes.c_int32)])


class PlcFrameConvDesc(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in ("N", "Cf", "H", "W", "Cout", "stride", "act")] +
                [("slope", ctypes.c_float), ("has_bias", ctypes.c_int32)])


_vp, _sz, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
_np = ctypes.POINTER(PlcConvNdDesc)
_fp = ctypes.POINTER(PlcFrameConvDesc)
_ip = ctypes.POINTER(ctypes.c_int)
_dp = ctypes.POINTER(PlcCellDesc)
_cp = ctypes.POINTER(PlcConvDesc)
_lp = ctypes.POINTER(PlcLossDesc)

# name -> (restype, argtypes); must list every symbol include/plc.h declares (tests check this)
SIGNATURES = {
    "plc_abi_version": (_int, []),
    "plc_last_error": (ctypes.c_char_p, []),
    "plc_device_supported": (_int, [_int]),
    "plc_packed_weight_bytes": (_sz, [_dp, _int]),
    "plc_pack_weight": (_int, [_dp, _int, _vp, _vp, _vp]),
    "plc_cell_fwd": (_int, [_dp] + [_vp] * 9),
    "plc_cell_fwd_zero_state_ok": (_int, [_dp]),
    "plc_bwd_workspace_bytes": (_sz, [_dp]),
    "plc_wgrad_acc_bytes": (_sz, [_dp]),
    "plc_wgrad_unpack": (_int, [_dp, _vp, _vp, _vp]),
    "plc_conv_wgrad_acc_bytes": (_sz, [_cp]),
    "plc_conv_wgrad_unpack": (_int, [_cp, _vp, _vp, _vp]),
    "plc_cell_bwd": (_int, [_dp] + [_vp] * 15 + [_sz, _vp]),
    "plc_cell_wgrad": (_int, [_dp] + [_vp] * 6),
    "plc_saved_gates_bytes": (_sz, [_dp]),
    "plc_cell_fwd_save": (_int, [_dp] + [_vp] * 9),
    "plc_cell_bwd_saved": (_int, [_dp] + [_vp] * 14 + [_sz, _vp]),
    "plc_conv_packed_weight_bytes": (_sz, [_cp, _int]),
    "plc_conv_pack_weight": (_int, [_cp, _int, _vp, _vp, _vp, _vp, _vp]),
    "plc_conv_fwd": (_int, [_cp, _vp, _vp, _vp, _vp, _vp]),
    "plc_conv_fwd_f32": (_int, [_cp, _vp, _vp, _vp, _vp, _vp]),
    "plc_conv_grad_mask": (_int, [_cp, _vp, _vp, _vp, _vp]),
    "plc_conv_im2col_narrow": (_int, [_cp, _int, _vp, _vp, _vp]),
    "plc_conv_bwd": (_int, [_cp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plc_convnd_out_shape": (_int, [_np, _ip, _ip, _ip]),
    "plc_convnd_packed_weight_bytes": (_sz, [_np, _int]),
    "plc_convnd_pack_weight": (_int, [_np, _int, _vp, _vp, _vp, _vp, _vp]),
    "plc_convnd_fwd": (_int, [_np, _vp, _vp, _vp, _vp, _vp]),
    "plc_convnd_grad_mask": (_int, [_np, _vp, _vp, _vp, _vp]),
    "plc_convnd_wgrad_acc_bytes": (_sz, [_np]),
    "plc_convnd_wgrad_unpack": (_int, [_np, _vp, _vp, _vp]),
    "plc_convnd_bwd": (_int, [_np, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plc_frameconv_out_shape": (_int, [_fp, _ip, _ip]),
    "plc_frameconv_fwd": (_int, [_fp, _vp, _vp, _vp, _vp, _vp]),
    "plc_frameconv_bwd": (_int, [_fp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plc_timing_enable": (_int, [_int]),
    "plc_timing_collect": (_int, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float),
                                  ctypes.POINTER(ctypes.c_double), _int]),
    "plc_debug_set_prof": (_int, [_vp]),
    "plc_debug_set_cta_group": (_int, [_int]),
    "plc_debug_set_patch": (_int, [_int]),
    "plc_nchw_f32_to_nhwc_bf16": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "plc_nhwc_bf16_to_nchw_f32": (_int, [_vp, _vp, _int, _int, _int, _int, _vp]),
    "plc_frontend_fwd": (_int, [_vp, _int, _int, _int, _int, _vp, _vp, _int, _int, _int, _vp, _vp]),
    "plc_frontend_tc_supported": (_int, [_int, _int, _int]),
    "plc_frontend_tc_fwd": (_int, [_vp, _int, _int, _int, _int, _int, _vp, _vp, _int, _vp, _vp]),
    "plc_head_fwd": (_int, [_vp, ctypes.c_long, _int, _vp, _vp, _int, _vp, _vp]),
    "plc_frames_to_nhwc": (_int, [_vp, _int, _int, _int, _int, _int, _int, _vp, _vp]),
    "plc_head_bwd": (_int, [_vp, ctypes.c_long, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plc_loss_workspace_bytes": (_sz, [_lp]),
    "plc_combined_loss": (_int, [_lp] + [_vp] * 9),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if needed and possible) libplc.so.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build`")
            from . import build as _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the binding drift apart
            fn.restype = res
            fn.argtypes = args
        if lib.plc_abi_version() != 1:
            raise RuntimeError("libplc.so ABI version mismatch")
        _lib = lib
        return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().plc_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


# ---- per-launch timing (include/plc.h: PlcKernelKind) ---------------------------------------------------------
KERNEL_KINDS = ("cell_fwd", "cell_fwd_zero", "bwd_gates", "bwd_dgrad", "bwd_wgrad", "conv_fwd", "conv_dgrad",
                "conv_wgrad", "frontend", "head", "loss", "pack", "elementwise")


def timing_enable(on: bool) -> None:
    """Start (clearing the record) / stop the library's per-launch CUDA-event timing.  Not for reported regions."""
    check(load().plc_timing_enable(1 if on else 0), "plc_timing_enable")


def timing_collect(capacity: int = 1 << 16):
    """-> list of (kind name, milliseconds, algorithmic FLOPs) in launch order; synchronises and clears the record."""
    kinds = (ctypes.c_int * capacity)()
    ms = (ctypes.c_float * capacity)()
    fl = (ctypes.c_double * capacity)()
    n = load().plc_timing_collect(kinds, ms, fl, capacity)
    return [(KERNEL_KINDS[kinds[i]], float(ms[i]), float(fl[i])) for i in range(min(n, capacity))]


# ---- packed-weight cache generation --------------------------------------------------------------------------
# Packed weight images are cached per module, keyed on the parameters' autograd version counters.  Fused optimizers
# (torch.optim.Adam(fused=True), ...) and writes through `.data` update parameters WITHOUT bumping that counter, so the
# key also carries a process-wide generation number that every optimizer step advances (global post-step hook below).
# Anything else that edits weights behind autograd's back must call `invalidate_packed_weights()`.
_generation = 0


def weight_generation() -> int:
    return _generation


def invalidate_packed_weights(*_args, **_kwargs) -> None:
    global _generation
    _generation += 1


def _install_optimizer_hook() -> None:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(invalidate_packed_weights)


_install_optimizer_hook()
