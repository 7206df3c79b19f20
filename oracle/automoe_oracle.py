"""ORACLE (test infrastructure, not product code) — a functional restatement of the
reference AutoMoE forward over a plain state_dict, in stock torch ops (fp32 by default).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product path (self-driving-model_b200/) never does.

Pinning: the reference ships no golden vectors for this path.  This
restatement is pinned against the reference itself: tests/golden/make_golden.py imports
/root/reference, loads the same synthetic state_dict into the reference modules and stores
the reference's outputs in tests/golden/*.npz; tests/test_oracle_golden.py checks this file
against those vectors.  Each function cites the reference lines it restates
(paths relative to the reference repo root).

Third-party arithmetic restated here: torchvision 0.26.0 resnet18 (BasicBlock.forward,
torchvision/models/resnet.py:89-105; ResNet._forward_impl :266-278) — expressed with
F.conv2d / F.batch_norm(eval) / F.relu / F.max_pool2d.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, never overridden by the reference
LN_EPS = 1e-5  # nn.LayerNorm default


def _bn(x, sd, p):
    """eval-mode BatchNorm2d: running statistics (reference runs inference under model.eval(),
    inference/run_automoe.py:155)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=BN_EPS)


def _linear(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(x, sd, p):
    w = sd[p + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[p + ".bias"], LN_EPS)


def basic_block(x, sd, p, stride):
    """torchvision BasicBlock.forward (resnet.py:89-105)."""
    identity = x
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
    out = F.relu(_bn(out, sd, p + ".bn1"))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = _bn(out, sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        identity = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0), sd, p + ".downsample.1")
    return F.relu(out + identity)


def resnet18_trunk(x, sd, p):
    """nn.Sequential(*list(resnet18.children())[:-2])  (models/experts/bdd_detection_expert.py:9-10;
    torchvision ResNet._forward_impl up to layer4)."""
    x = F.conv2d(x, sd[p + ".0.weight"], None, 2, 3)
    x = F.relu(_bn(x, sd, p + ".1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, stride in ((4, 1), (5, 2), (6, 2), (7, 2)):
        x = basic_block(x, sd, f"{p}.{li}.0", stride)
        x = basic_block(x, sd, f"{p}.{li}.1", 1)
    return x


def expert_head(feat, sd, p):
    """Conv2d(512,256,3,padding=1) -> ReLU -> Conv2d(256,N,1) (bdd_detection_expert.py:12-16,
    bdd_segmentation_expert.py:13-17, bdd_drivable_expert.py:13-17)."""
    h = F.relu(F.conv2d(feat, sd[p + ".0.weight"], sd[p + ".0.bias"], 1, 1))
    return F.conv2d(h, sd[p + ".2.weight"], sd[p + ".2.bias"])


def run_expert(x, sd, p, cfg):
    """BDDDetectionExpert.forward (bdd_detection_expert.py:18-24) /
    BDDSegmentationExpert.forward (bdd_segmentation_expert.py:19-23) /
    BDDDrivableExpert.forward (bdd_drivable_expert.py:19-23)."""
    feat = resnet18_trunk(x, sd, p + ".backbone")
    if cfg["type"] == "detection":
        nc = cfg.get("num_classes", 10)
        out = expert_head(feat, sd, p + ".head")
        return {"class_logits": out[:, :nc], "bbox_deltas": out[:, nc:]}
    low = expert_head(feat, sd, p + ".decoder")
    return F.interpolate(low, size=x.shape[-2:], mode="bilinear", align_corners=False)


def nuscenes_expert(image, sd, p, cfg):
    """NuScenesExpert.forward, image-only (models/experts/nuscenes_expert.py:150-190): resnet18 children()[:-1]
    (trunk + AdaptiveAvgPool2d(1)) -> Linear(512,256) -> + query_embed -> decoder (eval: Dropout = identity) -> heads."""
    feat = resnet18_trunk(image, sd, p + ".image_backbone").mean(dim=(2, 3))
    img = _linear(feat.to(sd[p + ".image_projection.weight"].dtype), sd, p + ".image_projection")      # [B,256]
    x = img.unsqueeze(1) + sd[p + ".query_embed.weight"].unsqueeze(0)                                   # [B,Q,256]
    x = F.relu(_linear(x, sd, p + ".decoder.0"))
    x = F.relu(_linear(x, sd, p + ".decoder.3"))
    return {"class_logits": _linear(x, sd, p + ".class_head"), "bbox_preds": _linear(x, sd, p + ".bbox_head")}


def nuscenes_extractor(expert_output, sd, p):
    """NuScenesExpertExtractor.forward (expert_extractors.py:125-137)."""
    flat = torch.cat([expert_output["class_logits"], expert_output["bbox_preds"]], dim=-1)
    flat = flat.reshape(flat.size(0), -1).to(sd[p + ".feature_extractor.0.weight"].dtype)
    v = F.relu(_linear(flat, sd, p + ".feature_extractor.0"))
    v = _linear(v, sd, p + ".feature_extractor.3")
    return _ln(v, sd, p + ".feature_extractor.4")


def extractor(expert_output, sd, p, cfg):
    """Detection/Segmentation/DrivableExpertExtractor.forward (expert_extractors.py:37-52,71-79,98-106):
    AdaptiveAvgPool2d(1) -> Flatten -> Linear -> ReLU -> Dropout(eval: id) -> Linear -> LayerNorm."""
    if cfg["type"] == "detection":
        expert_output = torch.cat([expert_output["class_logits"], expert_output["bbox_deltas"]], dim=1)
    v = expert_output.to(sd[p + ".feature_extractor.2.weight"].dtype).mean(dim=(2, 3))   # fp32 (fp64 in accuracy studies)
    v = F.relu(_linear(v, sd, p + ".feature_extractor.2"))
    v = _linear(v, sd, p + ".feature_extractor.5")
    return _ln(v, sd, p + ".feature_extractor.6")


def vehicle_state(batch):
    """AutoMoE._extract_context_features, simple branch (automoe.py:104-135): last column of
    [B,H] inputs, zeros for missing controls, cat -> [B,4]."""
    def last(t):
        if t.dim() == 2 and t.size(1) > 1:
            return t[:, -1:]
        if t.dim() > 2:
            return t.reshape(t.size(0), -1)[:, -1:]
        return t
    speed = batch["speed"]
    speed = speed[:, -1:] if (speed.dim() == 2 and speed.size(1) > 1) else speed
    if all(k in batch for k in ("speed", "steering", "throttle", "brake")):
        cols = [speed, last(batch["steering"]), last(batch["throttle"]), last(batch["brake"])]
    else:
        z = torch.zeros(speed.size(0), 1, device=speed.device)
        cols = [speed, z, z, z]
    return torch.cat([c.float() for c in cols], dim=-1)


def context_extractor(state, sd, p="context_extractor"):
    """SimpleContextExtractor.forward (context_features.py:151-165)."""
    v = F.relu(_linear(state.to(sd[p + ".encoder.0.weight"].dtype), sd, p + ".encoder.0"))
    v = _linear(v, sd, p + ".encoder.3")
    return _ln(v, sd, p + ".encoder.4")


def gating_network(features: List[torch.Tensor], context, sd, p="gating_network", temperature=1.0, context_only=False,
                   use_softmax=True):
    """GatingNetwork.forward (gating_network.py:122-175), eval mode, top_k=0 (all that is reachable through
    AutoMoE, automoe.py:83-91); softmax gate or, with use_softmax=False, the sigmoid gate of :159-160;
    context_only=True restates get_expert_weights (gating_network.py:177-199)."""
    c = F.relu(_linear(context, sd, p + ".context_encoder.context_encoder.0"))
    c = F.relu(_linear(c, sd, p + ".context_encoder.context_encoder.3"))
    E = len(features)
    processed = []
    for i, f in enumerate(features):
        if context_only:
            processed.append(torch.zeros(context.size(0), sd[p + ".output_projection.weight"].shape[0], device=context.device))
            continue
        q = p + f".expert_processors.{i}.processor"
        v = F.relu(_linear(f, sd, q + ".0"))
        v = _linear(v, sd, q + ".3")
        processed.append(_ln(v, sd, q + ".4"))
    gate_in = torch.cat([c] + processed, dim=1)
    g = F.relu(_linear(gate_in, sd, p + ".gate_network.0"))
    logits = _linear(g, sd, p + ".gate_network.3")
    if use_softmax:
        weights = F.softmax(logits / temperature, dim=1)
    else:
        weights = torch.sigmoid(logits)
        weights = weights / (weights.sum(dim=1, keepdim=True) + 1e-8)
    combined = torch.zeros_like(processed[0])
    for i in range(E):
        combined = combined + weights[:, i:i + 1] * processed[i]
    final = _linear(combined, sd, p + ".output_projection")
    return {"combined_output": final, "expert_weights": weights, "processed_expert_outputs": processed,
            "gate_logits": logits}


def policy_head(image, context, sd, p="policy_head", horizon=10):
    """EasyBackbone.forward + TrajectoryPolicy.forward (trajectory_head.py:27-33,55-63)."""
    x = image
    for ci, bi, k, pad in ((0, 1, 5, 2), (3, 4, 3, 1), (6, 7, 3, 1), (9, 10, 3, 1)):
        x = F.conv2d(x, sd[f"{p}.backbone.net.{ci}.weight"], sd[f"{p}.backbone.net.{ci}.bias"], 2, pad)
        x = F.relu(_bn(x, sd, f"{p}.backbone.net.{bi}"))
    feat = _linear(x.mean(dim=(2, 3)), sd, p + ".backbone.fc")
    v = torch.cat([feat, context], dim=1) if context is not None else feat
    outs = []
    for head in ("head_wp", "head_spd"):
        h = F.relu(_linear(v, sd, f"{p}.{head}.0"))
        h = F.relu(_linear(h, sd, f"{p}.{head}.2"))
        outs.append(_linear(h, sd, f"{p}.{head}.4"))
    return {"waypoints": outs[0].view(-1, horizon, 2), "speed": outs[1].view(-1, horizon)}


def automoe_forward(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], config: Dict) -> Dict:
    """AutoMoE.forward (automoe.py:189-233); experts of type detection / segmentation / drivable / nuscenes (image-only)."""
    image = batch["image"]
    ctx = context_extractor(vehicle_state(batch), sd)
    expert_outputs = [(nuscenes_expert if c["type"] == "nuscenes" else run_expert)(image, sd, f"experts.{i}", c)
                      for i, c in enumerate(config["experts"])]
    feats = [nuscenes_extractor(o, sd, f"expert_extractors.extractors.{i}") if c["type"] == "nuscenes"
             else extractor(o, sd, f"expert_extractors.extractors.{i}", c)
             for i, (o, c) in enumerate(zip(expert_outputs, config["experts"]))]
    g = gating_network(feats, ctx, sd, temperature=config["gating"].get("temperature", 1.0),
                       use_softmax=config["gating"].get("use_softmax", True))
    horizon = config["policy"].get("num_waypoints", 10)
    pol = policy_head(image, g["combined_output"], sd, horizon=horizon)
    speed_seq = pol["speed"]
    return {
        "waypoints": pol["waypoints"],
        "speed": speed_seq[:, -1:].contiguous(),
        "speed_seq": speed_seq,
        "expert_weights": g["expert_weights"],
        "expert_outputs": expert_outputs,
        "context_features": ctx,
        "combined_features": g["combined_output"],
        "gate_logits": g["gate_logits"],
        # extras for finer-grained parity checks (not in the reference dict)
        "_expert_features": feats,
        "_processed": g["processed_expert_outputs"],
    }


def get_expert_weights(sd, batch, config):
    """AutoMoE.get_expert_weights (automoe.py:235-238)."""
    ctx = context_extractor(vehicle_state(batch), sd)
    E = len(config["experts"])
    return gating_network([None] * E, ctx, sd, temperature=config["gating"].get("temperature", 1.0),
                          context_only=True, use_softmax=config["gating"].get("use_softmax", True))["expert_weights"]
