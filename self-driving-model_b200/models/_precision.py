"""Compute-precision selection shared by all modules.

precision = "bf16" | "fp32" | "auto".  "auto" follows the reference call sites: inside
torch.autocast (inference/run_automoe.py:51) the convolutions run in reduced precision
-> bf16 tcgen05 path; outside autocast the reference computes in fp32 -> fp32 kernels.
"""
import torch


def resolve_dtype(precision: str) -> torch.dtype:
    if precision == "bf16":
        return torch.bfloat16
    if precision == "fp32":
        return torch.float32
    if precision == "auto":
        return torch.bfloat16 if torch.is_autocast_enabled("cuda") else torch.float32
    raise ValueError(f"unknown precision {precision!r} (expected 'auto', 'bf16' or 'fp32')")
