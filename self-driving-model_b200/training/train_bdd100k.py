"""Detection-expert training step — drop-in for BDDTrainer._train_detection_batch
(training/train_bdd100k_ddp.py:117-186) on the sm_100a kernels.

Reference flow per batch: expert forward -> [B,Q,C]/[B,Q,4] -> HungarianMatcher -> Python loop scattering
the assignment into per-query targets -> CrossEntropy(ignore_index=num_classes) + 2 * SmoothL1(matched).
Here: differentiable expert forward (training/functional.py shims), the batched cost-matrix kernel + native
LSAP (training/hungarian_matcher.py), ONE scatter kernel for the whole batch and ONE fused loss kernel that
also produces the gradient w.r.t. the head output.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .._cabi import check, ctx, lib
from .._ops import ptr, stream_ptr
from .functional import _DetLoss
from .hungarian_matcher import HungarianMatcher


def xyxy_to_cxcywh(b: torch.Tensor) -> torch.Tensor:
    """torchvision.ops.box_convert(boxes, 'xyxy', 'cxcywh') (boxes.py): index arithmetic on [N,4]."""
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack(((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1), dim=-1)


def split_targets(gt_boxes: torch.Tensor, gt_labels: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
    """Padded [B,Nmax,4] xyxy / [B,Nmax] (-1 padding) -> list of ragged cxcywh targets (lines 122-148)."""
    out = []
    for b in range(gt_labels.size(0)):
        mask = gt_labels[b] != -1
        boxes = gt_boxes[b][mask]
        out.append({'boxes': xyxy_to_cxcywh(boxes) if boxes.numel() > 0 else boxes, 'labels': gt_labels[b][mask]})
    return out


def build_targets(indices: List[Tuple[torch.Tensor, torch.Tensor]], targets: List[Dict[str, torch.Tensor]], B: int, Q: int,
                  num_classes: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """target_classes [B*Q] (num_classes = unmatched) and target_boxes [B*Q,4] from the assignment: the
    reference's per-image indexing loop (lines 167-170) as one gather on the host side of the indices and
    one scatter kernel."""
    tcls = torch.full((B * Q,), num_classes, dtype=torch.int64, device=device)
    tbox = torch.zeros((B * Q, 4), dtype=torch.float32, device=device)
    pred_idx = torch.cat([p for p, _ in indices]) if indices else torch.zeros(0, dtype=torch.int64, device=device)
    n = int(pred_idx.numel())
    if n == 0:
        return tcls, tbox
    batch_of = torch.cat([torch.full((int(p.numel()),), b, dtype=torch.int32, device=device) for b, (p, _) in enumerate(indices)])
    labels = torch.cat([targets[b]['labels'][t].to(torch.int64) for b, (_, t) in enumerate(indices)])
    boxes = torch.cat([targets[b]['boxes'][t].to(torch.float32) for b, (_, t) in enumerate(indices)]).contiguous()
    check(lib().amoe_det_targets(ctx(device), ptr(pred_idx.contiguous()), ptr(batch_of), ptr(labels.contiguous()), ptr(boxes),
                                 n, Q, ptr(tcls), ptr(tbox), stream_ptr(device)), "det_targets")
    return tcls, tbox


def detection_losses(outputs: Dict[str, torch.Tensor], gt_boxes: torch.Tensor, gt_labels: torch.Tensor,
                     matcher: HungarianMatcher, num_classes: int, bbox_loss_weight: float = 2.0) -> Dict[str, torch.Tensor]:
    """total / class / bbox loss of _train_detection_batch for the expert outputs of one batch."""
    pred_logits, pred_boxes = outputs['class_logits'], outputs['bbox_deltas']
    B, C, H, W = pred_logits.shape
    Q = H * W
    head = outputs.get('_head_nhwc')
    if head is None:    # outputs of the inference path (no grad): NCHW -> NHWC copy
        head = torch.cat([pred_logits, pred_boxes], dim=1).permute(0, 2, 3, 1).contiguous().float()
    dev = head.device
    targets = split_targets(gt_boxes.to(dev), gt_labels.to(dev))
    with torch.no_grad():
        flat = head.detach().reshape(B, Q, C + 4)
        indices = matcher({'pred_logits': flat[..., :C].contiguous(), 'pred_boxes': flat[..., C:].contiguous()}, targets)
        tcls, tbox = build_targets(indices, targets, B, Q, num_classes, dev)
    losses = _DetLoss.apply(head, tcls, tbox, num_classes, bbox_loss_weight)
    report = losses.detach()     # only total_loss carries a gradient (the fused kernel differentiates the total)
    return {'total_loss': losses[0], 'class_loss': report[1], 'bbox_loss': report[2], 'num_matched': report[3],
            'indices': indices}


def train_detection_batch(model, batch: Dict[str, torch.Tensor], matcher: HungarianMatcher,
                          bbox_loss_weight: float = 2.0) -> torch.Tensor:
    """Same contract as BDDTrainer._train_detection_batch: returns the scalar total loss (autograd-connected
    to the expert's parameters); the caller runs backward / optimizer as the reference trainer does."""
    dev = next(model.parameters()).device
    images = batch['image'].to(dev)
    m = model.module if hasattr(model, 'module') else model
    outputs = model(images)
    return detection_losses(outputs, batch['bboxes'], batch['labels'], matcher, m.num_classes, bbox_loss_weight)['total_loss']
