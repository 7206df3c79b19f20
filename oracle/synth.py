"""ORACLE-side test infrastructure: deterministic synthetic weights and inputs
(SURVEY.md §8d): seeded CPU generators, BN running stats / affine parameters perturbed so
that BN folding is actually exercised (default init makes BN ~ identity)."""
from __future__ import annotations

from typing import Dict

import torch

CONFIG_3EXPERT = {
    "experts": [
        {"type": "detection", "num_classes": 10, "output_dim": 256, "pretrained_backbone": False},
        {"type": "segmentation", "num_classes": 19, "output_dim": 256, "pretrained_backbone": False},
        {"type": "drivable", "num_classes": 3, "output_dim": 256, "pretrained_backbone": False},
    ],
    "gating": {"processed_dim": 256, "hidden_dim": 128, "temperature": 1.0, "use_softmax": True},
    "context": {"type": "simple", "context_dim": 64},
    "policy": {"hidden_dim": 256, "num_waypoints": 10, "waypoint_dim": 2},
}


# the shipped models/configs/automoe/model_config.json: the three BDD experts + the image-only nuScenes expert
# (pretrained_backbone false: no network; gating keys top_k / noise_* are not forwarded by AutoMoE, automoe.py:83-91)
CONFIG_4EXPERT = {
    "experts": CONFIG_3EXPERT["experts"] + [
        {"type": "nuscenes", "num_queries": 196, "num_classes": 10, "output_dim": 256, "fusion": "sum", "use_lidar": False,
         "use_tnet": False, "bbox_dim": 4, "pretrained_backbone": False},
    ],
    "gating": dict(CONFIG_3EXPERT["gating"], top_k=2, noise_type="gumbel", noise_scale=0.0, apply_topk_at_eval=True),
    "context": CONFIG_3EXPERT["context"],
    "policy": CONFIG_3EXPERT["policy"],
}


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Fill a state_dict (shapes/keys taken from `template`) deterministically, key by key:
    conv: N(0, sqrt(2/fan_out)); linear weight: U(+-1/sqrt(fan_in)); biases: N(0,0.05);
    BN/LN weight: N(1,0.1), bias: N(0,0.1); running_mean: N(0,0.1); running_var: U(0.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, t in template.items():
        shape = tuple(t.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=t.dtype)
        elif k.endswith("running_mean"):
            out[k] = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            out[k] = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            out[k] = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
        elif len(shape) == 2:
            bound = 1.0 / shape[1] ** 0.5
            out[k] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif len(shape) == 1 and k.endswith("weight"):      # BN / LN scale
            out[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:                               # biases (conv, linear, BN, LN)
            out[k] = torch.randn(shape, generator=g) * (0.1 if _is_norm_bias(k, template) else 0.05)
        else:
            raise ValueError(f"unexpected tensor {k} {shape}")
    return out


def _is_norm_bias(k: str, template) -> bool:
    stem = k[: -len("bias")]
    return (stem + "running_mean") in template or (stem + "weight") in template and template[stem + "weight"].dim() == 1


def synth_batch(B: int, H: int = 256, W: int = 256, seed: int = 1, speed_seq: int = 1) -> Dict[str, torch.Tensor]:
    """image ~ N(0,1), speed ~ U(0,30), zero controls — as inference/run_automoe.py:41-49 builds them."""
    g = torch.Generator().manual_seed(seed)
    return {
        "image": torch.randn((B, 3, H, W), generator=g),
        "speed": torch.rand((B, speed_seq), generator=g) * 30.0,
        "steering": torch.zeros(B, 1),
        "throttle": torch.zeros(B, 1),
        "brake": torch.zeros(B, 1),
    }


def synth_matcher_case(B: int, Q: int, C: int, D: int, n_min: int, n_max: int, seed: int = 0):
    """Predictions like raw conv outputs but with positive sizes, ragged targets."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((B, Q, C), generator=g)
    boxes = torch.rand((B, Q, D), generator=g)
    if D == 4:
        boxes[..., 2:] = boxes[..., 2:] * 0.3 + 0.02
    elif D == 7:
        boxes[..., 3:6] = boxes[..., 3:6] * 0.3 + 0.02
    targets = []
    for b in range(B):
        n = int(torch.randint(n_min, n_max + 1, (1,), generator=g))
        tb = torch.rand((n, D), generator=g)
        if D == 4:
            tb[:, 2:] = tb[:, 2:] * 0.3 + 0.02
        elif D == 7:
            tb[:, 3:6] = tb[:, 3:6] * 0.3 + 0.02
        tl = torch.randint(0, C, (n,), generator=g)
        targets.append({"boxes": tb, "labels": tl})
    return {"pred_logits": logits, "pred_boxes": boxes}, targets


def synth_u8_frame(h, w, seed):
    """Seeded uint8 HWC frame: smooth gradient + noise + saturated patches (exercises clipping and flat areas)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255.0 / max(w - 1, 1)), (yy * 255.0 / max(h - 1, 1)), ((xx + yy) % 256)], -1)
    img = base + rng.randint(-60, 61, size=(h, w, 3))
    img[: h // 4, : w // 4] = 255
    img[-(h // 5):, -(w // 5):] = 0
    return np.clip(img, 0, 255).astype(np.uint8)
