"""SimpleContextExtractor — drop-in for models/context/context_features.py:137-165.
(The reference's "full" ContextFeatureExtractor is unreachable from the shipped config and
crashes on its own inputs; it is not reproduced.)"""
from typing import Dict

import torch
import torch.nn as nn

from ... import _ops
from .._gatepack import pack_gate_params, require_eval
from ..experts._trunk import params_stamp


class SimpleContextExtractor(nn.Module):
    """Simplified context extractor for basic vehicle state"""

    def __init__(self, context_dim: int = 64):
        super().__init__()
        self.context_dim = context_dim
        self.encoder = nn.Sequential(
            nn.Linear(4, 32),  # speed, steering, throttle, brake
            nn.ReLU(),
            nn.Dropout(0.1),
            nn.Linear(32, context_dim),
            nn.LayerNorm(context_dim),
        )
        self._flat = None

    def forward(self, speed, steering, throttle, brake) -> torch.Tensor:
        state = torch.cat([speed, steering, throttle, brake], dim=-1).float().contiguous()
        from .._train_forward import context_extractor_forward, wants_grad
        if wants_grad(self):
            return context_extractor_forward(self, state)
        stamp = (params_stamp([self]), state.device)
        if self._flat is None or self._flat[0] != stamp:
            self._flat = (stamp, pack_gate_params(self, None, None, [1], self.context_dim, 4, state.device))
        return _ops.gate(state, None, self._flat[1], [1], self.context_dim, 4, 1.0, mode=8)["context"]


def create_context_extractor(config: Dict) -> nn.Module:
    extractor_type = config.get('type', 'simple')
    if extractor_type == 'simple':
        return SimpleContextExtractor(context_dim=config.get('context_dim', 64))
    elif extractor_type == 'full':
        raise NotImplementedError("context type 'full' is dead code in the reference (see module docstring)")
    else:
        raise ValueError(f"Unknown context extractor type: {extractor_type}")
