"""Golden vectors of the gating/policy TRAINING step, produced by the UNMODIFIED
reference imported from /root/reference (CPU, fp32):

    python tests/golden/make_golden_train.py

The reference model (models/automoe.py) and the reference's own compute_gating_losses
(training/train_gating_network.py:21-74, loaded from its source file because the module's top-level
imports need tensorboard/dataloaders that are not part of the path) run one forward/backward on seeded
synthetic weights, inputs and targets.  Stored: the seven loss values and, for EVERY trainable
parameter, the gradient's L2 norm and its first 8 elements; full gradients of a few small tensors.
Three cases: eval-mode semantics (Dropout off, BatchNorm running statistics); train mode with Dropout p forced
to 0 and the frozen experts kept in eval() (policy BatchNorm on batch statistics); and the reference's actual
train_one_epoch state - model.train() on everything, so the frozen experts' BatchNorm layers also run on batch
statistics and update their running statistics (Dropout p forced to 0: its random stream is torch's own).
"""
import ast
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (used by the exec'd reference function)
import torchvision

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from models.automoe import create_automoe_model as ref_create  # noqa: E402  (reference)

from oracle import synth  # noqa: E402

OUT = Path(__file__).resolve().parent
VERS = np.array([f"torch {torch.__version__}", f"torchvision {torchvision.__version__}"])


def reference_loss_fn():
    """compute_gating_losses exactly as written in the reference file."""
    src = Path("/root/reference/training/train_gating_network.py").read_text()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_gating_losses"][0]
    ns = {"torch": torch, "F": F, "Dict": dict}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_gating_network.py", "exec"), ns)
    return ns["compute_gating_losses"]


def targets(B, horizon, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((B, horizon, 2), generator=g) * 5.0, torch.rand((B, horizon), generator=g) * 30.0


def case(name, B, H, train_mode, experts_eval=True):
    torch.manual_seed(0)
    ref = ref_create(synth.CONFIG_3EXPERT, "cpu")
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), 0), strict=True)
    ref.freeze_experts()
    if train_mode:
        ref.train()
        if experts_eval:
            ref.experts.eval()                   # frozen experts on running statistics (AutoMoE.frozen_experts_eval = True)
        # else: the reference's actual train_one_epoch state (train_gating_network.py:85): the frozen experts'
        # BatchNorm layers use batch statistics and update their running statistics
        for m in ref.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
    else:
        ref.eval()
    batch = synth.synth_batch(B, H, H, seed=3, speed_seq=1)
    wp, spd = targets(B, 10, 4)
    cfg = {}
    loss_fn = reference_loss_fn()
    pred = ref(batch)
    losses = loss_fn(pred, wp, spd, cfg)
    losses["total_loss"].backward()
    d = dict(versions=VERS, B=B, H=H, train_mode=train_mode, experts_eval=experts_eval,
             losses=np.array([losses[k].item() for k in ("total_loss", "ade", "fde", "speed", "smoothness", "load_balancing", "entropy")],
                             dtype=np.float64),
             waypoints=pred["waypoints"].detach().numpy(), expert_weights=pred["expert_weights"].detach().numpy())
    names, norms, heads = [], [], []
    for k, p in ref.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, k
        names.append(k)
        norms.append(p.grad.double().norm().item())
        h = np.zeros(8, dtype=np.float32)
        flat = p.grad.reshape(-1)[:8].numpy()
        h[:flat.size] = flat
        heads.append(h)
    d["grad_names"] = np.array(names)
    d["grad_norms"] = np.array(norms)
    d["grad_heads"] = np.stack(heads)
    for k in ("gating_network.gate_network.3.weight", "gating_network.gate_network.3.bias", "policy_head.head_spd.4.bias",
              "policy_head.backbone.net.1.weight", "policy_head.backbone.net.0.bias", "context_extractor.encoder.0.weight",
              "policy_head.backbone.net.0.weight", "expert_extractors.extractors.2.feature_extractor.2.weight"):
        d["full__" + k] = dict(ref.named_parameters())[k].grad.numpy()
    if train_mode:
        bn = ref.policy_head.backbone.net[1]
        d["bn1_running_mean"] = bn.running_mean.numpy()
        d["bn1_running_var"] = bn.running_var.numpy()
        if not experts_eval:
            sdr = ref.state_dict()
            for k in ("experts.0.backbone.1.running_mean", "experts.0.backbone.1.running_var", "experts.1.backbone.5.0.downsample.1.running_var",
                      "experts.2.backbone.7.1.bn2.running_mean", "experts.2.backbone.7.1.bn2.running_var"):
                d["stat__" + k] = sdr[k].numpy()
            d["expert_nbt"] = sdr["experts.1.backbone.4.0.bn1.num_batches_tracked"].numpy()
            d["seg_mean"] = pred["expert_outputs"][1].detach().mean(dim=(2, 3)).numpy()
            d["det_class_logits"] = pred["expert_outputs"][0]["class_logits"].detach().numpy()
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, "losses", d["losses"].round(5).tolist(), "n_trainable", len(names),
          "n_params", sum(p.numel() for p in ref.parameters() if p.requires_grad))


if __name__ == "__main__":
    case("train_eval_b4_64", 4, 64, False)
    case("train_trainmode_b4_64", 4, 64, True)
    case("train_refmode_b4_128", 4, 128, True, experts_eval=False)     # model.train() exactly as train_one_epoch leaves it
