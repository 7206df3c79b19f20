"""ORACLE (test infrastructure, not product code) — numpy restatement of
HungarianMatcher.forward (training/hungarian_matcher.py:20-85).

Third-party arithmetic restated: torchvision 0.26.0 ops.boxes.box_convert (cxcywh->xyxy,
_box_convert.py: x1 = cx - 0.5*w ...), generalized_box_iou (boxes.py:374-401:
iou - (hull - union)/hull), torch.cdist(p=1), softmax; the assignment itself is
scipy.optimize.linear_sum_assignment (scipy 1.18.1), exactly the reference's call
(hungarian_matcher.py:79).  All float32, one rounded op per step, like the eager reference.

Pinned against the reference by tests/golden/make_golden.py (matcher_*.npz).
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

f32 = np.float32


def softmax_rows(x):
    x = x.astype(f32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=f32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)


def cxcywh_to_xyxy(b):
    cx, cy, w, h = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    half = f32(0.5)
    return np.stack([cx - half * w, cy - half * h, cx + half * w, cy + half * h], axis=1).astype(f32)


def bev_xyxy(b):
    """[cx,cy,cz,w,l,h,yaw] -> axis-aligned BEV corners (hungarian_matcher.py:56-64)."""
    xc, yc, w, l = b[:, 0], b[:, 1], b[:, 3], b[:, 4]
    two = f32(2)
    return np.stack([xc - w / two, yc - l / two, xc + w / two, yc + l / two], axis=1).astype(f32)


def generalized_box_iou(a, b):
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / union
        lti = np.minimum(a[:, None, :2], b[None, :, :2])
        rbi = np.maximum(a[:, None, 2:], b[None, :, 2:])
        whi = np.clip(rbi - lti, 0, None)
        areai = whi[..., 0] * whi[..., 1]
        return (iou - (areai - union) / areai).astype(f32)


def cost_matrix(logits, boxes, tgt_boxes, tgt_labels, w_class=1.0, w_bbox=5.0, w_giou=2.0):
    """One image: logits [Q,C], boxes [Q,D], tgt_boxes [N,D], tgt_labels [N] -> [Q,N] float32."""
    logits, boxes, tgt_boxes = logits.astype(f32), boxes.astype(f32), tgt_boxes.astype(f32)
    D = boxes.shape[1]
    prob = softmax_rows(logits)
    c_class = -prob[:, tgt_labels]
    diff = np.abs(boxes[:, None, :] - tgt_boxes[None, :, :])
    c_bbox = np.zeros(diff.shape[:2], f32)
    for k in range(D):  # sequential sum like the cdist kernel
        c_bbox = c_bbox + diff[:, :, k]
    if w_giou > 0 and D == 4:
        c_giou = -generalized_box_iou(cxcywh_to_xyxy(boxes), cxcywh_to_xyxy(tgt_boxes))
    elif w_giou > 0 and D == 7:
        c_giou = -generalized_box_iou(bev_xyxy(boxes), bev_xyxy(tgt_boxes))
    else:
        c_giou = np.zeros_like(c_bbox)
    return (f32(w_bbox) * c_bbox + f32(w_class) * c_class + f32(w_giou) * c_giou).astype(f32)


def match(logits, boxes, targets, w_class=1.0, w_bbox=5.0, w_giou=2.0):
    """Batched: logits [B,Q,C], boxes [B,Q,D], targets = list of (boxes [Ni,D], labels [Ni]).
    Returns list of (row_idx, col_idx) int64 arrays and the list of cost matrices."""
    out, costs = [], []
    for b in range(logits.shape[0]):
        tb, tl = targets[b]
        C = cost_matrix(logits[b], boxes[b], tb, tl, w_class, w_bbox, w_giou)
        r, c = linear_sum_assignment(C)
        out.append((r.astype(np.int64), c.astype(np.int64)))
        costs.append(C)
    return out, costs
