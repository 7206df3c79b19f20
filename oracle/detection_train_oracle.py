"""ORACLE (test infrastructure, not product code) — the reference's detection-expert training step
(BDDTrainer._train_detection_batch, training/train_bdd100k_ddp.py:117-186) restated in stock torch ops
over a plain state_dict; differentiable through torch.autograd.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.  Pinned by
tests/golden/make_golden_det.py, which executes the reference's own method on the reference expert.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import automoe_oracle as O
from . import matcher_oracle as MO


def _bn_train(x, sd, p):
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=O.BN_EPS)


def trunk_train(x, sd, p, batch_stats=True):
    """resnet18 children()[:-2] with train-mode BatchNorm (torchvision resnet.py:89-105,266-278)."""
    bn = _bn_train if batch_stats else O._bn
    x = F.conv2d(x, sd[p + ".0.weight"], None, 2, 3)
    x = F.relu(bn(x, sd, p + ".1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, stride in ((4, 1), (5, 2), (6, 2), (7, 2)):
        for bi, s in ((0, stride), (1, 1)):
            q = f"{p}.{li}.{bi}"
            idn = x
            out = F.relu(bn(F.conv2d(x, sd[q + ".conv1.weight"], None, s, 1), sd, q + ".bn1"))
            out = bn(F.conv2d(out, sd[q + ".conv2.weight"], None, 1, 1), sd, q + ".bn2")
            if (q + ".downsample.0.weight") in sd:
                idn = bn(F.conv2d(x, sd[q + ".downsample.0.weight"], None, s, 0), sd, q + ".downsample.1")
            x = F.relu(out + idn)
    return x


def detection_forward_train(x, sd, num_classes=10, batch_stats=True):
    """BDDDetectionExpert.forward (bdd_detection_expert.py:18-24) in train mode."""
    out = O.expert_head(trunk_train(x, sd, "backbone", batch_stats), sd, "head")
    return {"class_logits": out[:, :num_classes], "bbox_deltas": out[:, num_classes:]}


def xyxy_to_cxcywh(b):
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack(((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1), dim=-1)


def detection_loss(outputs: Dict[str, torch.Tensor], gt_boxes, gt_labels, num_classes=10, bbox_loss_weight=2.0):
    """train_bdd100k_ddp.py:122-186 (targets, reshape, matcher, scatter, CE(ignore) + w * SmoothL1)."""
    pl, pb = outputs["class_logits"], outputs["bbox_deltas"]
    B, C, H, W = pl.shape
    Q = H * W
    pl = pl.permute(0, 2, 3, 1).reshape(B, Q, C)                                     # :140
    pb = pb.permute(0, 2, 3, 1).reshape(B, Q, 4)                                     # :141
    targets = []
    for b in range(B):
        mask = gt_labels[b] != -1                                                    # :124
        boxes = gt_boxes[b][mask]
        targets.append((xyxy_to_cxcywh(boxes) if boxes.numel() > 0 else boxes, gt_labels[b][mask]))   # :144-152
    idx, _ = MO.match(pl.detach().cpu().numpy(), pb.detach().cpu().numpy(),
                      [(tb.cpu().numpy(), tl.cpu().numpy()) for tb, tl in targets])   # :154 HungarianMatcher
    tcls = torch.full((B * Q,), num_classes, dtype=torch.int64, device=pl.device)    # :163
    tbox = torch.zeros((B * Q, 4), dtype=pb.dtype, device=pl.device)                 # :166
    for b, (pi, ti) in enumerate(idx):                                               # :168-170
        pi, ti = torch.as_tensor(pi, device=pl.device), torch.as_tensor(ti, device=pl.device)
        tcls[b * Q + pi] = targets[b][1][ti]
        tbox[b * Q + pi] = targets[b][0][ti].to(pb.dtype)
    class_loss = F.cross_entropy(pl.reshape(B * Q, C), tcls, ignore_index=num_classes)   # :172 (ctor :50)
    matched = tcls != num_classes
    if matched.any():
        bbox_loss = F.smooth_l1_loss(pb.reshape(B * Q, 4)[matched], tbox[matched], reduction="mean")  # :178 (ctor :51)
    else:
        bbox_loss = torch.tensor(0.0, device=pl.device)
    return {"total_loss": class_loss + bbox_loss_weight * bbox_loss, "class_loss": class_loss, "bbox_loss": bbox_loss,
            "indices": idx}


def synth_detection_batch(B, H, W, n_max, seed):
    """images ~ N(0,1); per image U{1..n_max} boxes (pixel xyxy inside the frame... normalised to [0,1] here,
    as the cxcywh predictions are raw conv outputs of order 1), labels U{0..9}, padded with -1."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn((B, 3, H, W), generator=g)
    boxes = torch.full((B, n_max, 4), -1.0)
    labels = torch.full((B, n_max), -1, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(1, n_max + 1, (1,), generator=g))
        xy = torch.rand((n, 2), generator=g) * 0.6
        wh = torch.rand((n, 2), generator=g) * 0.35 + 0.03
        boxes[b, :n] = torch.cat([xy, xy + wh], dim=1)
        labels[b, :n] = torch.randint(0, 10, (n,), generator=g)
    return {"image": images, "bboxes": boxes, "labels": labels}
