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


class NuScenesExpertExtractor(ExpertOutputExtractor):
    """Extracts features from nuScenes expert outputs (expert_extractors.py:108-137): cat(class_logits, bbox_preds) ->
    flatten [B, Q*(C+bbox_dim)] -> Linear(.,512)+ReLU+Dropout -> Linear(512,out) -> LayerNorm.  The 2,744-wide input does not
    fit the fused gate kernel's shared-memory rows, so this extractor runs as two GEMM launches + one LayerNorm launch and
    hands its feature to the gate kernel (amoe_gate_fwd_ex2, n_ch = 0)."""

    def __init__(self, output_dim: int = 256, num_queries: int = 100, num_classes: int = 10, bbox_dim: int = 7):
        super().__init__(output_dim)
        self.num_queries = num_queries
        self.num_classes = num_classes
        self.bbox_dim = bbox_dim
        self.feature_extractor = nn.Sequential(
            nn.Linear(num_queries * (num_classes + self.bbox_dim), 512),
            nn.ReLU(),
            nn.Dropout(0.1),
            nn.Linear(512, output_dim),
            nn.LayerNorm(output_dim),
        )

    def in_channels(self) -> int:
        return 0            # not a pooled-channel extractor: its feature reaches the gate kernel ready-made

    def forward(self, expert_output: Dict[str, torch.Tensor]) -> torch.Tensor:
        from ...training import functional as TF
        flat = expert_output.get('_flat')
        if flat is None:
            flat = torch.cat([expert_output['class_logits'], expert_output['bbox_preds']], dim=-1)
        flat = flat.reshape(flat.size(0), -1).float()
        if not flat.is_cuda:
            raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
        fe = self.feature_extractor
        p = float(fe[2].p) if self.training else 0.0
        h = TF.linear(flat, fe[0], relu=True, drop_p=p)
        h = TF.linear(h, fe[3])
        return TF.layer_norm(h, fe[4])


class ExpertOutputManager(nn.Module):
    """Manages multiple expert output extractors as a registered module (expert_extractors.py:139-157)"""

    def __init__(self, extractors: List[ExpertOutputExtractor]):
        super().__init__()
        self.extractors = nn.ModuleList(extractors)

    def extract_features(self, expert_outputs: List) -> List[torch.Tensor]:
        return [extractor(out) for extractor, out in zip(self.extractors, expert_outputs)]


def create_expert_extractors(expert_configs: List[Dict]) -> ExpertOutputManager:
    """Same factory as expert_extractors.py:159-200."""
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
            extractor = NuScenesExpertExtractor(output_dim=config.get('output_dim', 256),
                                                num_queries=config.get('num_queries', 100),
                                                num_classes=config.get('num_classes', 10),
                                                bbox_dim=config.get('bbox_dim', 7))
        else:
            raise ValueError(f"Unknown expert type: {expert_type}")
        extractors.append(extractor)
    return ExpertOutputManager(extractors)
