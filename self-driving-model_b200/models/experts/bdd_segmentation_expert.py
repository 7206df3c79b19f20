"""BDDSegmentationExpert — drop-in for models/experts/bdd_segmentation_expert.py:5-23."""
from ... import _ops
from ._base import BDDExpertBase
from ._trunk import make_head, make_resnet18_trunk


class BDDSegmentationExpert(BDDExpertBase):
    head_attr = "decoder"
    upsample_to_input = True

    def __init__(self, num_classes=19, pretrained_backbone=True):
        super().__init__()
        self.num_classes = num_classes
        self.backbone = make_resnet18_trunk(pretrained_backbone)
        self.decoder = make_head(num_classes)

    def format_output(self, low, H, W, dtype):
        # F.interpolate(logits_lowres, size=x.shape[-2:], mode="bilinear", align_corners=False)
        return _ops.upsample_bilinear_nchw(low, H, W, dtype)
