from .bdd_detection_expert import BDDDetectionExpert
from .bdd_drivable_expert import BDDDrivableExpert
from .bdd_segmentation_expert import BDDSegmentationExpert
from .nuscenes_expert import NuScenesExpert

__all__ = ["BDDDetectionExpert", "BDDDrivableExpert", "BDDSegmentationExpert", "NuScenesExpert"]
