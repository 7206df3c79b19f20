"""Differentiable forward of the trainable part of AutoMoE (context extractor, expert extractors,
GatingNetwork, TrajectoryPolicy) for training/train_gating_network.py.

Taken whenever a module is in train mode, or in eval mode with autograd recording and trainable
parameters (eval-mode semantics - Dropout off, BatchNorm running statistics - stay differentiable,
which is what the gradient-parity tests compare against the reference).  Every layer is one of the
autograd shims in training/functional.py, i.e. sm_100a kernels forward and backward, fp32.

Reference lines restated: context_features.py:143-165, expert_extractors.py:27-52,
gating_network.py:12-20,37-43,94-99,122-175, trajectory_head.py:8-33,44-63.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import _ops
from ..training import functional as TF


def wants_grad(module: nn.Module) -> bool:
    """True when the forward of `module` must build an autograd graph / use train-mode semantics."""
    if module.training:
        return True
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


def _p(drop: nn.Module, training: bool) -> float:
    return float(drop.p) if (training and isinstance(drop, nn.Dropout)) else 0.0


def context_extractor_forward(mod, state: torch.Tensor) -> torch.Tensor:
    """SimpleContextExtractor: Linear(4,32)+ReLU+Dropout -> Linear(32,ctx) -> LayerNorm."""
    enc, tr = mod.encoder, mod.training
    h = TF.linear(state, enc[0], relu=True, drop_p=_p(enc[2], tr))
    h = TF.linear(h, enc[3])
    return TF.layer_norm(h, enc[4])


def extractor_forward(mod, pooled: torch.Tensor) -> torch.Tensor:
    """*ExpertExtractor after the average pool: Linear(C,512)+ReLU+Dropout -> Linear(512,out) -> LayerNorm."""
    fe, tr = mod.feature_extractor, mod.training
    h = TF.linear(pooled, fe[2], relu=True, drop_p=_p(fe[4], tr))
    h = TF.linear(h, fe[5])
    return TF.layer_norm(h, fe[6])


def gating_forward(gn, features: List[torch.Tensor], context: torch.Tensor) -> Dict[str, torch.Tensor]:
    """GatingNetwork.forward (softmax gate, no top-k)."""
    tr = gn.training
    ce = gn.context_encoder.context_encoder
    c = TF.linear(context, ce[0], relu=True, drop_p=_p(ce[2], tr))
    c = TF.linear(c, ce[3], relu=True, drop_p=_p(ce[5], tr))
    processed = []
    for f, proc in zip(features, gn.expert_processors):
        pr = proc.processor
        v = TF.linear(f, pr[0], relu=True, drop_p=_p(pr[2], tr))
        v = TF.linear(v, pr[3])
        processed.append(TF.layer_norm(v, pr[4]))
    gate_in = torch.cat([c] + processed, dim=1)                       # plumbing: one copy
    g = TF.linear(gate_in, gn.gate_network[0], relu=True, drop_p=_p(gn.gate_network[2], tr))
    logits = TF.linear(g, gn.gate_network[3])
    weights, combined = TF.gate_combine(logits, processed, gn.temperature, gn.use_softmax)
    final = TF.linear(combined, gn.output_projection)
    return {'combined_output': final, 'expert_weights': weights, 'processed_expert_outputs': processed,
            'gate_logits': logits}


def policy_forward(pol, image: torch.Tensor, context: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """EasyBackbone (4 x conv s2 + BatchNorm + ReLU, GAP, fc) + both TrajectoryPolicy heads."""
    if not image.is_cuda:
        raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
    x = TF.image_nhwc4(image)                                         # [B,H,W,4], zero 4th channel
    net = pol.backbone.net
    for ci, bi in ((0, 1), (3, 4), (6, 7), (9, 10)):
        x = TF.conv_bn_relu(x, net[ci], net[bi], batch_stats=net[bi].training)
    feat = TF.linear(TF.global_avg_pool(x), pol.backbone.fc)
    if context is not None:
        if pol.context_dim == 0:
            raise ValueError("TrajectoryPolicy was built with context_dim=0 but a context was given")
        v = torch.cat([feat, context.float()], dim=1)
    else:
        if pol.context_dim != 0:
            raise ValueError("TrajectoryPolicy expects a context of dim %d" % pol.context_dim)
        v = feat
    outs = []
    for head in (pol.head_wp, pol.head_spd):
        h = TF.linear(v, head[0], relu=True)
        h = TF.linear(h, head[2], relu=True)
        outs.append(TF.linear(h, head[4]))
    return {"waypoints": outs[0].view(-1, pol.horizon, 2), "speed": outs[1].view(-1, pol.horizon)}
