"""ORACLE (test infrastructure, not product code) — the reference's gating losses and training
semantics restated in stock torch ops, differentiable through torch.autograd.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
Pinned against the reference itself by tests/golden/make_golden_train.py (loss values and gradients
of the unmodified reference on seeded inputs, stored in tests/golden/train_*.npz).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import automoe_oracle as O


def compute_gating_losses(pred: Dict[str, torch.Tensor], target_wp, target_spd, config: Dict) -> Dict[str, torch.Tensor]:
    """training/train_gating_network.py:21-74."""
    ade = F.l1_loss(pred["waypoints"], target_wp)                                     # :26
    fde = F.l1_loss(pred["waypoints"][:, -1, :], target_wp[:, -1, :])                 # :27
    pred_spd = pred.get("speed_seq", pred.get("speed"))                              # :28
    if pred_spd is not None and pred_spd.dim() == 2 and target_spd.dim() == 2 and pred_spd.size(1) == target_spd.size(1):
        speed_loss = F.l1_loss(pred_spd, target_spd)                                  # :30
    else:
        pred_last = pred.get("speed")
        if pred_last is not None and pred_last.dim() == 2 and pred_last.size(1) == 1:
            speed_loss = F.l1_loss(pred_last, target_spd[:, -1:].contiguous())         # :33-35
        else:
            speed_loss = torch.zeros((), device=target_spd.device)
    d = pred["waypoints"][:, 1:, :] - pred["waypoints"][:, :-1, :]                    # :39
    smooth = F.l1_loss(d[:, 1:, :], d[:, :-1, :])                                     # :40
    w = pred["expert_weights"]
    if config.get('use_load_balancing', True):
        mean_usage = w.mean(dim=0)                                                    # :45
        lb = F.mse_loss(mean_usage, torch.ones_like(mean_usage) / mean_usage.size(0)) # :46-47
    else:
        lb = torch.tensor(0.0, device=w.device)
    if config.get('use_entropy_loss', True):
        entropy = -(w * torch.log(w + 1e-8)).sum(dim=1).mean()                        # :52
        ent = -entropy                                                                # :53
    else:
        ent = torch.tensor(0.0, device=w.device)
    total = (config.get('ade_weight', 1.0) * ade + config.get('fde_weight', 2.0) * fde +
             config.get('speed_weight', 0.2) * speed_loss + config.get('smoothness_weight', 0.1) * smooth +
             config.get('load_balancing_weight', 0.01) * lb + config.get('entropy_weight', 0.001) * ent)   # :57-64
    return {"total_loss": total, "ade": ade, "fde": fde, "speed": speed_loss, "smoothness": smooth,
            "load_balancing": lb, "entropy": ent}


TRAINABLE_PREFIXES = ("context_extractor.", "expert_extractors.", "gating_network.", "policy_head.")


def is_trainable_key(k: str) -> bool:
    """Keys AdamW touches after freeze_experts() (automoe.py:269-273): everything but experts.*,
    and only floating-point parameters (BatchNorm buffers are not parameters)."""
    return k.startswith(TRAINABLE_PREFIXES) and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))


def policy_head_batchstats(image, context, sd, p="policy_head", horizon=10):
    """EasyBackbone with train-mode BatchNorm (batch statistics), otherwise O.policy_head."""
    x = image
    for ci, bi, k, pad in ((0, 1, 5, 2), (3, 4, 3, 1), (6, 7, 3, 1), (9, 10, 3, 1)):
        x = F.conv2d(x, sd[f"{p}.backbone.net.{ci}.weight"], sd[f"{p}.backbone.net.{ci}.bias"], 2, pad)
        x = F.relu(F.batch_norm(x, None, None, sd[f"{p}.backbone.net.{bi}.weight"], sd[f"{p}.backbone.net.{bi}.bias"],
                                training=True, eps=O.BN_EPS))
    feat = O._linear(x.mean(dim=(2, 3)), sd, p + ".backbone.fc")
    v = torch.cat([feat, context], dim=1) if context is not None else feat
    outs = []
    for head in ("head_wp", "head_spd"):
        h = F.relu(O._linear(v, sd, f"{p}.{head}.0"))
        h = F.relu(O._linear(h, sd, f"{p}.{head}.2"))
        outs.append(O._linear(h, sd, f"{p}.{head}.4"))
    return {"waypoints": outs[0].view(-1, horizon, 2), "speed": outs[1].view(-1, horizon)}


def run_expert_batchstats(x, sd, p, cfg):
    """An expert inside model.train() (train_gating_network.py:85): torchvision BasicBlocks with BATCH statistics
    (running statistics are updated as a side effect in the reference; the outputs do not depend on them)."""
    from .detection_train_oracle import trunk_train
    feat = trunk_train(x, sd, p + ".backbone", batch_stats=True)
    if cfg["type"] == "detection":
        nc = cfg.get("num_classes", 10)
        out = O.expert_head(feat, sd, p + ".head")
        return {"class_logits": out[:, :nc], "bbox_deltas": out[:, nc:]}
    low = O.expert_head(feat, sd, p + ".decoder")
    return F.interpolate(low, size=x.shape[-2:], mode="bilinear", align_corners=False)


def training_forward(sd, batch, config, policy_batch_stats: bool = False, expert_batch_stats: bool = False):
    """AutoMoE.forward as train_one_epoch sees it, Dropout off, experts frozen; gradients flow into the
    TRAINABLE_PREFIXES parameters only.  expert_batch_stats=True is the reference's actual state after model.train()
    (frozen experts' BatchNorm on batch statistics); False = after model.experts.eval() (running statistics)."""
    image = batch["image"]
    ctx = O.context_extractor(O.vehicle_state(batch), sd)
    run = run_expert_batchstats if expert_batch_stats else O.run_expert
    with torch.no_grad():
        expert_outputs = [run(image, sd, f"experts.{i}", c) for i, c in enumerate(config["experts"])]
    feats = [O.extractor(o, sd, f"expert_extractors.extractors.{i}", c)
             for i, (o, c) in enumerate(zip(expert_outputs, config["experts"]))]
    g = O.gating_network(feats, ctx, sd, temperature=config["gating"].get("temperature", 1.0))
    horizon = config["policy"].get("num_waypoints", 10)
    pol = (policy_head_batchstats if policy_batch_stats else O.policy_head)(image, g["combined_output"], sd, horizon=horizon)
    return {"waypoints": pol["waypoints"], "speed": pol["speed"][:, -1:].contiguous(), "speed_seq": pol["speed"],
            "expert_weights": g["expert_weights"], "gate_logits": g["gate_logits"], "combined_features": g["combined_output"],
            "expert_outputs": expert_outputs}
