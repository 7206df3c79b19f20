"""Compute-precision selection shared by all modules.

precision = "bf16" | "fp32" | "auto".  "auto" follows the reference call sites: inside
torch.autocast (inference/run_automoe.py:51) the convolutions run in reduced precision
-> bf16 tcgen05 path; outside autocast the reference computes in fp32 -> fp32 kernels.

The reference's `torch.autocast('cuda', enabled=True)` defaults to float16; the tensor-core kernels here
compute in bfloat16 with fp32 accumulation whatever 16-bit type autocast names (same 1e-2 tolerance class
against fp32, wider exponent range; outputs of the small heads stay fp32).  A float16 request is honoured with
bf16 and reported once, not silently.
"""
import warnings

import torch

_warned_fp16 = False


def resolve_dtype(precision: str) -> torch.dtype:
    global _warned_fp16
    if precision == "bf16":
        return torch.bfloat16
    if precision == "fp32":
        return torch.float32
    if precision == "auto":
        if not torch.is_autocast_enabled("cuda"):
            return torch.float32
        if not _warned_fp16 and torch.get_autocast_dtype("cuda") == torch.float16:
            _warned_fp16 = True
            warnings.warn("automoe_b200: torch.autocast requested float16; the sm_100a kernels compute in bfloat16 with "
                          "fp32 accumulation instead (pass dtype=torch.bfloat16 to autocast to silence this)")
        return torch.bfloat16
    raise ValueError(f"unknown precision {precision!r} (expected 'auto', 'bf16' or 'fp32')")
