"""ctypes binding of include/automoe_b200.h (libautomoe_b200.so).

There is no CPU fallback: if the shared library is missing or a call fails, this
module raises.  Tensors cross the boundary as raw device pointers + sizes and the
current CUDA stream; PyTorch owns every buffer.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libautomoe_b200.so"

F32, BF16 = 0, 1

_lib = None
_lock = threading.RLock()  # re-entrant: ctx() loads the library while holding it
_ctxs: dict[int, C.c_void_p] = {}

_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/automoe_b200.h one to one
SIGNATURES = {
    "amoe_abi_version": (_I, []),
    "amoe_last_error": (C.c_char_p, []),
    "amoe_create": (_I, [_I, C.POINTER(_P)]),
    "amoe_destroy": (_I, [_P]),
    "amoe_sm_count": (_I, [_P]),
    "amoe_launch_count": (_L, [_P]),
    "amoe_set_walk_reverse": (_I, [_P, _I]),
    "amoe_image_nchw_to_nhwc": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_image_nchw_to_nhwc_padded": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_image_nchw_to_nhwc_padded_v": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
    "amoe_stage_u8_hwc_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _F, _P]),
    "amoe_normalize_u8_hwc_to_nchw_fwd": (_I, [_P, _P, _P, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _P]),
    "amoe_resample_u8_fwd": (_I, [_P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _P]),
    "amoe_stem_pool_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, C.POINTER(_P), C.POINTER(_I), _P]),
    "amoe_stem_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, C.POINTER(_P), C.POINTER(_I), _P]),
    "amoe_stem_fwd_f32tc_supported": (_I, [_I] * 4),
    "amoe_stem_split_frame": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "amoe_stem_fwd_f32tc": (_I, [_P] * 6 + [_I] * 7 + [_P]),
    "amoe_conv2d_rowwin_fwd": (_I, [_P] * 6 + [_I] * 13 + [_P]),
    "amoe_pack_conv_weight": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_fold_bn": (_I, [_P, _P, _P, _P, _P, _F, _P, _I, _P, _P, _P]),
    "amoe_conv2d_fwd": (_I, [_P] * 7 + [_I] * 20 + [_P]),
    "amoe_conv2d_dual_fwd": (_I, [_P] * 10 + [_I] * 16 + [_P]),
    "amoe_split3_bf16": (_I, [_P, _P, _P, _L, _I, _P]),
    "amoe_pack_conv_weight_split6": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "amoe_conv2d_f32tc_supported": (_I, [_I] * 7),
    "amoe_conv2d_fwd_f32tc": (_I, [_P] * 6 + [_I] * 12 + [_P]),
    "amoe_conv2d_fwd_f32tc_grouped": (_I, [_P] * 7 + [_I] * 14 + [_P]),
    "amoe_conv2d_bwd_data_f32tc": (_I, [_P] * 6 + [_I] * 11 + [_P]),
    "amoe_split3_padded": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "amoe_conv3x3_wgrad_f32tc_supported": (_I, [_I, _I]),
    "amoe_conv3x3_wgrad_f32tc_workspace_floats": (_L, [_P, _I, _I, _L]),
    "amoe_conv3x3_wgrad_f32tc": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _L, _P]),
    "amoe_conv3x3_flat_fwd": (_I, [_P] * 7 + [_I] * 7 + [_P]),
    "amoe_conv3x3_flat_fwd_strided": (_I, [_P] * 7 + [_I] * 7 + [C.c_int64, C.c_int64, _P]),
    "amoe_conv3x3_flat_supported": (_I, [_I] * 4),
    "amoe_conv2d_tc_supported": (_I, [_I] * 6),
    "amoe_maxpool3x3s2_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_head1x1_pool_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_upsample_bilinear_nchw_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_mean_hw_nchw_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "amoe_gate_fwd": (_I, [_P, _P, _P, _P, _L, _I, _I, C.POINTER(_I), _I, _I, _F, _I] + [_P] * 7),
    "amoe_gate_fwd_ex": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, C.POINTER(_I), _I, _I, _F, _I] + [_P] * 7),
    "amoe_gate_fwd_ex2": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, C.POINTER(_I), _I, _I, _F, _I, _P] + [_P] * 7),
    "amoe_bcast_add_relu": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "amoe_policy_head_fwd": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "amoe_policy_head_fwd_ex": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "amoe_mean_hw_nhwc_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "amoe_hungarian_cost_fwd": (_I, [_P] * 7 + [_I] * 5 + [_F] * 3 + [_P]),
    "amoe_lsap_batched_host": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I]),
    # training step (gating + policy)
    "amoe_linear_fwd": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, C.c_uint64, _P]),
    "amoe_linear_fwd_dseed": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, C.c_uint64, _P, _P]),
    "amoe_linear_bwd": (_I, [_P, _P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _F, _P]),
    "amoe_layernorm_fwd": (_I, [_P] * 7 + [_I, _I, _F, _P]),
    "amoe_layernorm_bwd": (_I, [_P] * 9 + [_I, _I, _P]),
    "amoe_gate_combine_fwd": (_I, [_P, _P, _P, _L, _I, _F, _P, _P, _I, _I, _I, _P]),
    "amoe_gate_combine_fwd_ex": (_I, [_P, _P, _P, _L, _I, _F, _I, _P, _P, _I, _I, _I, _P]),
    "amoe_gate_combine_bwd_ex": (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _P, _P, _P, _L, _I, _I, _I, _P]),
    "amoe_gate_combine_bwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _P, _P, _L, _I, _I, _I, _P]),
    "amoe_gating_loss_fwd_bwd": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, C.POINTER(_F), _I, _I, _P, _P, _P, _P, _P]),
    "amoe_sq_norm": (_I, [_P, _P, _L, _P, _I, _P, _P]),
    "amoe_fused_clip_adamw": (_I, [_P, _P, _P, _P, _P, _L, _P] + [_F] * 7 + [_I, _P]),
    "amoe_fused_clip_adamw_dstep": (_I, [_P, _P, _P, _P, _P, _L, _P] + [_F] * 7 + [_P, _P]),
    "amoe_train_tick": (_I, [_P, _P, _P, _P]),
    "amoe_colreduce_workspace_floats": (_L, [_L, _I]),
    "amoe_bn_train_fwd": (_I, [_P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _L, _I, _I, _P]),
    "amoe_bn_train_fwd_grouped": (_I, [_P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _I, _L, _I, _I, _P]),
    "amoe_bn_apply_fwd": (_I, [_P] * 7 + [_L, _I, _I, _P]),
    "amoe_bn_bwd": (_I, [_P] * 11 + [_L, _I, _I, _P]),
    "amoe_colsum": (_I, [_P, _P, _P, _P, _L, _I, _F, _P]),
    "amoe_conv2d_bwd_data": (_I, [_P] * 4 + [_I] * 13 + [_P]),
    "amoe_conv2d_bwd_weight_workspace_floats": (_L, [_P] + [_I] * 7),
    "amoe_conv2d_bwd_weight": (_I, [_P] * 5 + [_L] + [_I] * 13 + [_P]),
    "amoe_gap_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "amoe_gap_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    # detection-expert training step
    "amoe_maxpool3x3s2_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "amoe_maxpool3x3s2_bwd_ws": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "amoe_add_relu_fwd": (_I, [_P, _P, _P, _P, _L, _P]),
    "amoe_relu_bwd": (_I, [_P, _P, _P, _P, _L, _P]),
    "amoe_upsample_bilinear_nchw_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "amoe_det_targets": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "amoe_det_loss_fwd_bwd": (_I, [_P, _P, _I, _P, _I, _P, _P, _L, _I, _I, _F, _P, _P, _I, _P, _I, _P]),
}


class AmoeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not LIB_PATH.exists():
                    raise AmoeError(
                        f"{LIB_PATH} not found: build it with `python self-driving-model_b200/build.py` "
                        "(or __graft_entry__.build()); there is no fallback path")
                l = C.CDLL(str(LIB_PATH))
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().amoe_last_error().decode("utf-8", "replace")
        raise AmoeError(f"{what or 'amoe call'} failed ({rc}): {msg}")


def ctx(device: torch.device | int | None = None) -> C.c_void_p:
    """Per-device context handle (created lazily; needs a CUDA device of sm_100)."""
    if device is None:
        idx = torch.cuda.current_device()
    elif isinstance(device, int):
        idx = device
    else:
        device = torch.device(device)
        if device.type != "cuda":
            raise AmoeError(f"automoe_b200 kernels need a CUDA device, got {device}; there is no CPU path")
        idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _ctxs.get(idx)
    if h is None:
        lib()
        with _lock:
            h = _ctxs.get(idx)
            if h is None:
                torch.cuda.init()
                out = _P()
                with torch.cuda.device(idx):
                    check(lib().amoe_create(idx, C.byref(out)), "amoe_create")
                _ctxs[idx] = h = out
    return h


def stream_ptr(device=None) -> C.c_void_p:
    return _P(torch.cuda.current_stream(device).cuda_stream)


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    return _P(0) if t is None else _P(t.data_ptr())


def launch_count(device=None) -> int:
    return int(lib().amoe_launch_count(ctx(device)))


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise AmoeError(f"unsupported dtype {dt}")
