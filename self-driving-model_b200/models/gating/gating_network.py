"""GatingNetwork — drop-in for models/gating/gating_network.py (ContextEncoder,
ExpertOutputProcessor, GatingNetwork.forward / get_expert_weights / get_gating_logits).

The noisy top-k branch of the reference is never enabled through AutoMoE
(models/automoe.py:83-91 does not forward top_k), so top_k > 0 is rejected here.
"""
from typing import Dict, List

import torch
import torch.nn as nn

from ... import _ops
from .._gatepack import pack_gate_params, require_eval
from ..experts._trunk import params_stamp


class ContextEncoder(nn.Module):
    """Encodes driving context to determine expert weights (gating_network.py:6-29)"""

    def __init__(self, context_dim: int = 64, hidden_dim: int = 128):
        super().__init__()
        self.context_dim = context_dim
        self.hidden_dim = hidden_dim
        self.context_encoder = nn.Sequential(
            nn.Linear(context_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.1),
        )


class ExpertOutputProcessor(nn.Module):
    """Processes and normalizes expert outputs for gating (gating_network.py:31-53)"""

    def __init__(self, expert_output_dim: int, processed_dim: int = 256):
        super().__init__()
        self.expert_output_dim = expert_output_dim
        self.processed_dim = processed_dim
        self.processor = nn.Sequential(
            nn.Linear(expert_output_dim, processed_dim), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(processed_dim, processed_dim), nn.LayerNorm(processed_dim),
        )


class GatingNetwork(nn.Module):
    """Mixture of Experts gating network"""

    def __init__(self, num_experts: int, context_dim: int = 64, expert_output_dims: List[int] = None,
                 processed_dim: int = 256, hidden_dim: int = 128, temperature: float = 1.0,
                 use_softmax: bool = True, top_k: int = 0, noise_type: str = 'gumbel',
                 noise_scale: float = 1.0, apply_topk_at_eval: bool = False):
        super().__init__()
        self.num_experts = num_experts
        self.context_dim = context_dim
        self.processed_dim = processed_dim
        self.hidden_dim = hidden_dim
        self.temperature = temperature
        self.use_softmax = use_softmax
        self.top_k = max(0, int(top_k))
        self.noise_type = noise_type
        self.noise_scale = float(noise_scale)
        self.apply_topk_at_eval = bool(apply_topk_at_eval)
        if self.top_k > 0:
            raise NotImplementedError("noisy top-k gating is unreachable through AutoMoE and not implemented")
        if expert_output_dims is None:
            expert_output_dims = [256] * num_experts
        self.context_encoder = ContextEncoder(context_dim, hidden_dim)
        self.expert_processors = nn.ModuleList([ExpertOutputProcessor(dim, processed_dim) for dim in expert_output_dims])
        self.gate_network = nn.Sequential(
            nn.Linear(hidden_dim + processed_dim * num_experts, hidden_dim), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim, num_experts),
        )
        self.output_projection = nn.Linear(processed_dim, processed_dim)
        self._flat = None

    def _gate_kind(self) -> int:
        """Mode bit of the fused kernel: 32 = sigmoid gate (use_softmax=False, reference gating_network.py:159-160)."""
        return 0 if self.use_softmax else 32

    def _params(self, device):
        stamp = (params_stamp([self]), device)
        if self._flat is None or self._flat[0] != stamp:
            self._flat = (stamp, pack_gate_params(None, None, self, [1] * self.num_experts, self.context_dim,
                                                  self.hidden_dim, device))
        return self._flat[1]

    def forward(self, expert_outputs: List[torch.Tensor], context: torch.Tensor) -> Dict[str, torch.Tensor]:
        from .._train_forward import gating_forward, wants_grad
        if wants_grad(self):
            return gating_forward(self, [t.float() for t in expert_outputs], context.float())
        E = self.num_experts
        feats = torch.stack([t.float() for t in expert_outputs], dim=0).contiguous()  # [E,B,256]
        context = context.float().contiguous()
        out = _ops.gate(context, feats, self._params(context.device), [1] * E, self.context_dim, self.hidden_dim,
                        self.temperature, mode=2 | 4 | self._gate_kind())
        return {
            'combined_output': out['combined'],
            'expert_weights': out['weights'],
            'processed_expert_outputs': [out['processed'][e] for e in range(E)],
            'gate_logits': out['gate_logits'],
        }

    def _context_only(self, context: torch.Tensor):
        require_eval(self, "GatingNetwork")
        context = context.float().contiguous()
        return _ops.gate(context, None, self._params(context.device), [1] * self.num_experts, self.context_dim,
                         self.hidden_dim, self.temperature, mode=1 | 2 | self._gate_kind())

    def get_expert_weights(self, context: torch.Tensor) -> torch.Tensor:
        """Expert weights from the context alone (gating_network.py:177-199)."""
        return self._context_only(context)['weights']

    def get_gating_logits(self, context: torch.Tensor) -> torch.Tensor:
        """Raw gating logits, context-only path (gating_network.py:201-207)."""
        return self._context_only(context)['gate_logits']
