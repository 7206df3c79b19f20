"""Flat fp32 parameter buffer of the fused gate kernel (layout documented in csrc/gate.cu)."""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _ops

EXT_HID = 512   # hidden width of the expert extractors (expert_extractors.py:30,64,91)
FEAT = 256      # expert feature dim == processed dim the kernel is built for


def _z(*shape):
    return torch.zeros(*shape)


def gate_param_tensors(context_extractor, extractors: Optional[list], gating, n_ch: List[int], ctx_dim: int,
                       hidden: int) -> List[torch.Tensor]:
    """Tensors in kernel order; a missing module contributes zeros of the right size."""
    E = len(n_ch)
    t: List[torch.Tensor] = []
    if context_extractor is not None:
        enc = context_extractor.encoder
        t += [enc[0].weight, enc[0].bias, enc[3].weight, enc[3].bias, enc[4].weight, enc[4].bias]
    else:
        t += [_z(32, 4), _z(32), _z(ctx_dim, 32), _z(ctx_dim), _z(ctx_dim), _z(ctx_dim)]
    for e in range(E):
        if n_ch[e] == 0:
            # feature supplied from outside the kernel (amoe_gate_fwd_ex2): zero-width first layer, the rest unused
            t += [_z(EXT_HID, 0), _z(EXT_HID), _z(FEAT, EXT_HID), _z(FEAT), _z(FEAT), _z(FEAT)]
        elif extractors is not None:
            lin = [m for m in extractors[e].feature_extractor if isinstance(m, torch.nn.Linear)]
            ln = [m for m in extractors[e].feature_extractor if isinstance(m, torch.nn.LayerNorm)][0]
            if lin[0].weight.shape != (EXT_HID, n_ch[e]) or lin[1].weight.shape != (FEAT, EXT_HID):
                raise NotImplementedError("fused gate kernel is built for Linear(C,512)->Linear(512,256) extractors")
            t += [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, ln.weight, ln.bias]
        else:
            t += [_z(EXT_HID, n_ch[e]), _z(EXT_HID), _z(FEAT, EXT_HID), _z(FEAT), _z(FEAT), _z(FEAT)]
    if gating is not None:
        ce = gating.context_encoder.context_encoder
        t += [ce[0].weight, ce[0].bias, ce[3].weight, ce[3].bias]
        for e in range(E):
            pr = gating.expert_processors[e].processor
            if pr[0].weight.shape != (FEAT, FEAT):
                raise NotImplementedError("fused gate kernel is built for processed_dim == output_dim == 256")
            t += [pr[0].weight, pr[0].bias, pr[3].weight, pr[3].bias, pr[4].weight, pr[4].bias]
        gn = gating.gate_network
        t += [gn[0].weight, gn[0].bias, gn[3].weight, gn[3].bias,
              gating.output_projection.weight, gating.output_projection.bias]
    else:
        t += [_z(hidden, ctx_dim), _z(hidden), _z(hidden, hidden), _z(hidden)]
        for e in range(E):
            t += [_z(FEAT, FEAT), _z(FEAT), _z(FEAT, FEAT), _z(FEAT), _z(FEAT), _z(FEAT)]
        t += [_z(hidden, hidden + FEAT * E), _z(hidden), _z(E, hidden), _z(E), _z(FEAT, FEAT), _z(FEAT)]
    return t


def pack_gate_params(context_extractor, extractors, gating, n_ch, ctx_dim, hidden, device) -> torch.Tensor:
    return _ops.flat_params(gate_param_tensors(context_extractor, extractors, gating, n_ch, ctx_dim, hidden), device)


def require_eval(module, what: str):
    if module.training:
        raise NotImplementedError(
            f"{what}: train-mode forward (active Dropout, autograd) is not implemented in the sm_100a path yet; "
            "call .eval()")
