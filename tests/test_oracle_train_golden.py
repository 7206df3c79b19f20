"""CPU: the training oracle (oracle/gating_train_oracle.py) reproduces the loss values and gradients
the unmodified reference produced (tests/golden/train_*.npz, made by make_golden_train.py)."""
import numpy as np
import pytest
import torch

from _util import rel_err
from oracle import gating_train_oracle as GT
from oracle import synth


def _targets(B, horizon, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((B, horizon, 2), generator=g) * 5.0, torch.rand((B, horizon), generator=g) * 30.0


def oracle_step(sd, batch, wp, spd, policy_batch_stats, expert_batch_stats=False):
    sd = {k: v.clone() for k, v in sd.items()}
    for k, v in sd.items():
        if GT.is_trainable_key(k) and v.is_floating_point():
            v.requires_grad_(True)
    pred = GT.training_forward(sd, batch, synth.CONFIG_3EXPERT, policy_batch_stats=policy_batch_stats,
                               expert_batch_stats=expert_batch_stats)
    losses = GT.compute_gating_losses(pred, wp, spd, {})
    losses["total_loss"].backward()
    return sd, pred, losses


@pytest.mark.parametrize("name", ["train_eval_b4_64", "train_trainmode_b4_64", "train_refmode_b4_128"])
def test_training_oracle_matches_reference(name, golden_dir):
    from automoe_b200.models.automoe import create_automoe_model
    g = np.load(golden_dir / f"{name}.npz")
    B, H, train_mode = int(g["B"]), int(g["H"]), bool(g["train_mode"])
    template = create_automoe_model(synth.CONFIG_3EXPERT, "cpu").state_dict()
    sd0 = synth.synth_state_dict(template, 0)
    batch = synth.synth_batch(B, H, H, seed=3, speed_seq=1)
    wp, spd = _targets(B, 10, 4)
    experts_eval = bool(g["experts_eval"]) if "experts_eval" in g.files else True
    sd, pred, losses = oracle_step(sd0, batch, wp, spd, train_mode, expert_batch_stats=not experts_eval)
    if not experts_eval:
        assert rel_err(pred["expert_outputs"][1].mean(dim=(2, 3)), g["seg_mean"]) < 1e-5
        assert rel_err(pred["expert_outputs"][0]["class_logits"], g["det_class_logits"]) < 1e-5
    got = np.array([losses[k].item() for k in ("total_loss", "ade", "fde", "speed", "smoothness", "load_balancing", "entropy")])
    assert np.allclose(got, g["losses"], rtol=2e-6, atol=1e-7), (got, g["losses"])
    assert rel_err(pred["waypoints"].detach(), g["waypoints"]) < 1e-5
    names = [str(n) for n in g["grad_names"]]
    assert len(names) == 82 and sum(sd[n].numel() for n in names) == 2870657
    for n, norm, head in zip(names, g["grad_norms"], g["grad_heads"]):
        gr = sd[n].grad
        assert gr is not None, n
        assert abs(gr.double().norm().item() - norm) <= 2e-4 * max(norm, 1e-6) + 1e-9, (n, gr.norm().item(), norm)
        k = min(8, gr.numel())
        assert np.allclose(gr.reshape(-1)[:k].numpy(), head[:k], rtol=2e-3, atol=1e-6 + 1e-4 * float(np.abs(head).max())), n
    for key in g.files:
        if key.startswith("full__"):
            assert rel_err(sd[key[6:]].grad, g[key]) < 2e-4, key


def _det_expert_sd():
    from automoe_b200.models.automoe import create_automoe_model
    full = synth.synth_state_dict(create_automoe_model(synth.CONFIG_3EXPERT, "cpu").state_dict(), 0)
    return {k[len("experts.0."):]: v for k, v in full.items() if k.startswith("experts.0.")}


def test_detection_training_oracle_matches_reference(golden_dir):
    """oracle/detection_train_oracle.py against the reference's BDDDetectionExpert + _train_detection_batch."""
    from oracle import detection_train_oracle as DO
    g = np.load(golden_dir / "det_train_b3_128x160.npz")
    sd = {k: v.clone() for k, v in _det_expert_sd().items()}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    batch = DO.synth_detection_batch(int(g["B"]), int(g["H"]), int(g["W"]), int(g["n_max"]), int(g["seed"]))
    out = DO.detection_forward_train(batch["image"], sd)
    losses = DO.detection_loss(out, batch["bboxes"], batch["labels"])
    losses["total_loss"].backward()
    assert abs(losses["total_loss"].item() - float(g["total_loss"])) < 2e-5 * float(g["total_loss"])
    for n, norm in zip([str(x) for x in g["grad_names"]], g["grad_norms"]):
        got = sd[n].grad.double().norm().item()
        assert abs(got - norm) <= 1e-3 * max(norm, 1e-4) + 1e-7, (n, got, norm)
    for key in g.files:
        if key.startswith("full__") and float(np.abs(g[key]).max()) > 1e-6:
            assert rel_err(sd[key[6:]].grad, g[key]) < 1e-3, key
