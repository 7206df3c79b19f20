"""Expert output extractors — same classes/keys as models/experts/expert_extractors.py.

Inside AutoMoE.forward these parameters are consumed by the fused gate kernel, which starts
from the experts' pooled low-res logits.  Called stand-alone (as the reference's unit tests
do) an extractor pools its input with the NCHW mean kernel and runs the same fused kernel
restricted to its own MLP + LayerNorm.
"""
from typing import Dict, List

import torch
import torch.nn as nn

from ... import _ops
from .._gatepack import pack_gate_params, require_eval
from ._trunk import params_stamp


class ExpertOutputExtractor(nn.Module):
    """Base class for extracting features from expert outputs"""

    def __init__(self, output_dim: int = 256):
        super().__init__()
        self.output_dim = output_dim
        self._flat = None

    def _make_mlp(self, in_ch: int, output_dim: int) -> nn.Sequential:
        return nn.Sequential(
            nn.AdaptiveAvgPool2d((1, 1)),
            nn.Flatten(),
            nn.Linear(in_ch, 512),
            nn.ReLU(),
            nn.Dropout(0.1),
            nn.Linear(512, output_dim),
            nn.LayerNorm(output_dim),
        )

    def in_channels(self) -> int:
        return self.feature_extractor[2].weight.shape[1]

    def _as_map(self, expert_output) -> torch.Tensor:
        return expert_output

    def forward(self, expert_output) -> torch.Tensor:
        x = self._as_map(expert_output)
        pooled = _ops.mean_hw_nchw(x)  # AdaptiveAvgPool2d((1,1)) + Flatten
        from .._train_forward import extractor_forward, wants_grad
        if wants_grad(self):   # experts are frozen upstream: the pooled logits are constants of the graph
            return extractor_forward(self, pooled)
        n_ch = [self.in_channels()]
        stamp = (params_stamp([self]), x.device)
        if self._flat is None or self._flat[0] != stamp:
            self._flat = (stamp, pack_gate_params(None, [self], None, n_ch, 4, 4, x.device))
        B = pooled.shape[0]
        dummy_ctx = torch.zeros((B, 4), device=x.device, dtype=torch.float32)
        out = _ops.gate(dummy_ctx, pooled, self._flat[1], n_ch, 4, 4, 1.0, mode=2 | 16)
        return out["features"][0]


class DetectionExpertExtractor(ExpertOutputExtractor):
    """Extracts features from detection expert outputs (expert_extractors.py:20-52)"""

    def __init__(self, output_dim: int = 256, num_classes: int = 10):
        super().__init__(output_dim)
        self.num_classes = num_classes
        self.feature_extractor = self._make_mlp(num_classes + 4, output_dim)

    def _as_map(self, expert_output: Dict[str, torch.Tensor]) -> torch.Tensor:
        return torch.cat([expert_output['class_logits'], expert_output['bbox_deltas']], dim=1)


class SegmentationExpertExtractor(ExpertOutputExtractor):
    """Extracts features from segmentation expert outputs (expert_extractors.py:54-79)"""

    def __init__(self, output_dim: int = 256, num_classes: int = 19):
        super().__init__(output_dim)
        self.num_classes = num_classes
        self.feature_extractor = self._make_mlp(num_classes, output_dim)


class DrivableExpertExtractor(ExpertOutputExtractor):
    """Extracts features from drivable area expert outputs (expert_extractors.py:81-106)"""

    def __init__(self, output_dim: int = 256, num_classes: int = 3):
        super().__init__(output_dim)
        self.num_classes = num_classes
        self.feature_extractor = self._make_mlp(num_classes, output_dim)


class ExpertOutputManager(nn.Module):
    """Manages multiple expert output extractors as a registered module (expert_extractors.py:139-157)"""

    def __init__(self, extractors: List[ExpertOutputExtractor]):
        super().__init__()
        self.extractors = nn.ModuleList(extractors)

    def extract_features(self, expert_outputs: List) -> List[torch.Tensor]:
        return [extractor(out) for extractor, out in zip(self.extractors, expert_outputs)]


def create_expert_extractors(expert_configs: List[Dict]) -> ExpertOutputManager:
    """Same factory as expert_extractors.py:159-200 (nuScenes expert: see DESIGN.md, out of scope)."""
    extractors = []
    for config in expert_configs:
        expert_type = config['type']
        if expert_type == 'detection':
            extractor = DetectionExpertExtractor(output_dim=config.get('output_dim', 256),
                                                 num_classes=config.get('num_classes', 10))
        elif expert_type == 'segmentation':
            extractor = SegmentationExpertExtractor(output_dim=config.get('output_dim', 256),
                                                    num_classes=config.get('num_classes', 19))
        elif expert_type == 'drivable':
            extractor = DrivableExpertExtractor(output_dim=config.get('output_dim', 256),
                                                num_classes=config.get('num_classes', 3))
        elif expert_type == 'nuscenes':
            raise NotImplementedError("the nuScenes expert is not part of the B200 hot path yet (SURVEY.md §8f)")
        else:
            raise ValueError(f"Unknown expert type: {expert_type}")
        extractors.append(extractor)
    return ExpertOutputManager(extractors)
