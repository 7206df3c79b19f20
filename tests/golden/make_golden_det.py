"""Golden vectors of the detection-expert TRAINING step, produced by the UNMODIFIED
reference imported from /root/reference (CPU, fp32):

    python tests/golden/make_golden_det.py

Runs the reference's BDDDetectionExpert (models/experts/bdd_detection_expert.py), its HungarianMatcher and
the body of BDDTrainer._train_detection_batch (training/train_bdd100k_ddp.py:117-186, executed from its own
source with a stand-in `self`, because the trainer module's imports need tensorboard / datasets) in train
mode on seeded weights and inputs; stores losses, the assignment, gradient norms / heads of every parameter,
a few full gradients and updated BatchNorm running statistics.
"""
import ast
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torchvision
from torchvision.ops import box_convert

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from models.experts.bdd_detection_expert import BDDDetectionExpert as RefDet  # noqa: E402  (reference)
from training.hungarian_matcher import HungarianMatcher as RefMatcher  # noqa: E402  (reference)

from oracle import detection_train_oracle as DO  # noqa: E402
from oracle import synth  # noqa: E402

OUT = Path(__file__).resolve().parent
VERS = np.array([f"torch {torch.__version__}", f"torchvision {torchvision.__version__}"])


def reference_method():
    src = Path("/root/reference/training/train_bdd100k_ddp.py").read_text()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BDDTrainer"][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_train_detection_batch"][0]
    ns = {"torch": torch, "box_convert": box_convert}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_bdd100k_ddp.py", "exec"), ns)
    return ns["_train_detection_batch"]


def expert_state_dict(template):
    """experts.0.* of the synthetic AutoMoE state dict = the detection expert's own state dict."""
    from automoe_b200.models.automoe import create_automoe_model
    full = synth.synth_state_dict(create_automoe_model(synth.CONFIG_3EXPERT, "cpu").state_dict(), 0)
    sd = {k[len("experts.0."):]: v for k, v in full.items() if k.startswith("experts.0.")}
    assert set(sd) == set(template), set(template) ^ set(sd)
    return sd


def case(name, B, H, W, n_max, seed):
    ref = RefDet(num_classes=10, pretrained_backbone=False)
    ref.load_state_dict(expert_state_dict(ref.state_dict()), strict=True)
    ref.train()
    batch = DO.synth_detection_batch(B, H, W, n_max, seed)
    me = types.SimpleNamespace(device=torch.device("cpu"), model=ref, matcher=RefMatcher(),
                               class_loss_fn=nn.CrossEntropyLoss(ignore_index=ref.num_classes),
                               bbox_loss_fn=nn.SmoothL1Loss(reduction="mean"), config={})
    loss = reference_method()(me, batch)
    loss.backward()
    d = dict(versions=VERS, B=B, H=H, W=W, n_max=n_max, seed=seed, total_loss=np.float64(loss.item()))
    names, norms, heads = [], [], []
    for k, p in ref.named_parameters():
        names.append(k)
        norms.append(p.grad.double().norm().item())
        h = np.zeros(8, dtype=np.float32)
        flat = p.grad.reshape(-1)[:8].numpy()
        h[:flat.size] = flat
        heads.append(h)
    d["grad_names"], d["grad_norms"], d["grad_heads"] = np.array(names), np.array(norms), np.stack(heads)
    params = dict(ref.named_parameters())
    for k in ("head.2.weight", "head.2.bias", "head.0.bias", "backbone.0.weight", "backbone.1.weight", "backbone.7.1.bn2.bias",
              "backbone.5.0.downsample.0.weight"):
        d["full__" + k] = params[k].grad.numpy()
    d["bn1_running_mean"] = ref.backbone[1].running_mean.numpy()
    d["bn1_running_var"] = ref.backbone[1].running_var.numpy()
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, "loss", loss.item(), "params", sum(p.numel() for p in ref.parameters()))


if __name__ == "__main__":
    case("det_train_b3_128x160", 3, 128, 160, 5, 31)
