"""Shared helpers for the parity tests."""
import numpy as np
import torch

from oracle import synth


def rel_err(a, b):
    """max|a-b| / max|b|  (the 'relative' of BASELINE.md §5)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def golden_config(g):
    """Model config of a golden automoe case (the sigmoid-gate case stores use_softmax=False)."""
    cfg = dict(synth.CONFIG_4EXPERT if ("four_experts" in g.files and bool(g["four_experts"])) else synth.CONFIG_3EXPERT)
    if "use_softmax" in g.files:
        cfg["gating"] = dict(cfg["gating"], use_softmax=bool(g["use_softmax"]))
    return cfg


def build_b200_model(device, precision="auto", seed=0, config=None):
    from automoe_b200.models.automoe import create_automoe_model
    cfg = dict(config or synth.CONFIG_3EXPERT)
    cfg["precision"] = precision
    m = create_automoe_model(cfg, "cpu")
    sd = synth.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd, strict=True)
    m = m.to(device).eval()
    m.device = device
    return m, sd


def golden_batch(g):
    """Re-create the inputs of a golden automoe case from its seeds."""
    B, H, W, sq = int(g["B"]), int(g["H"]), int(g["W"]), int(g["speed_seq"])
    batch = synth.synth_batch(B, H, W, seed=1, speed_seq=sq)
    if bool(g["controls_seq"]):
        gen = torch.Generator().manual_seed(7)
        for k in ("steering", "throttle", "brake"):
            batch[k] = torch.rand((B, sq), generator=gen) - 0.5
    return batch
