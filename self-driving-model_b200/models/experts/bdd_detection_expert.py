"""BDDDetectionExpert — same constructor, attributes, state_dict keys and outputs as the
reference (models/experts/bdd_detection_expert.py:4-31); forward runs on sm_100a kernels."""
import torch

from ... import _ops
from ._base import BDDExpertBase
from ._trunk import make_head, make_resnet18_trunk


class BDDDetectionExpert(BDDExpertBase):
    head_attr = "head"
    upsample_to_input = False

    def __init__(self, num_classes=10, pretrained_backbone=True):
        super().__init__()
        self.num_classes = num_classes
        self.backbone = make_resnet18_trunk(pretrained_backbone)
        self.head = make_head(num_classes + 4)

    def format_output(self, low, H, W, dtype):
        # low: [B,h,w,C+4] fp32 NHWC -> NCHW [B,C+4,h,w]; the two outputs are channel slices
        # of one tensor, as in the reference (bdd_detection_expert.py:21-24)
        out = _ops.upsample_bilinear_nchw(low, low.shape[1], low.shape[2], dtype)
        return {
            "class_logits": out[:, :self.num_classes, :, :],
            "bbox_deltas": out[:, self.num_classes:, :, :],
        }

    def format_output_train(self, low, H=None, W=None):
        out = low.permute(0, 3, 1, 2)       # NCHW view of the NHWC head output (no copy)
        return {
            "class_logits": out[:, :self.num_classes, :, :],
            "bbox_deltas": out[:, self.num_classes:, :, :],
            "_head_nhwc": low,              # extra key: lets the fused detection loss skip the permutes
        }

    def predict(self, x):
        output = self.forward(x)
        return {
            "class_probs": output["class_logits"].float().softmax(dim=1),
            "bbox_deltas": output["bbox_deltas"].float().sigmoid(),
        }
