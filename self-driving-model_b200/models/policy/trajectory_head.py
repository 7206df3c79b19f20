"""TrajectoryPolicy / EasyBackbone — drop-in for models/policy/trajectory_head.py:5-63.

Four stride-2 conv+BN+ReLU stages run through amoe_conv2d_fwd (BN folded into the
epilogue); global-average-pool, fc, concat with the gated context and both MLP heads
run in one fused kernel (csrc/policy.cu).
"""
from typing import Dict, Optional

import torch
import torch.nn as nn

from ... import _ops
from .._gatepack import require_eval
from .._precision import resolve_dtype
from ..experts._trunk import ParamHolder, params_stamp, stage_image


class EasyBackbone(nn.Module):
    def __init__(self, in_channels: int = 3, out_dim: int = 512):
        super().__init__()
        self.net = ParamHolder(
            nn.Conv2d(in_channels, 32, kernel_size=5, stride=2, padding=2), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
            nn.Conv2d(32, 64, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
            nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
            nn.Conv2d(128, 256, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
        )
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(256, out_dim)


class TrajectoryPolicy(nn.Module):
    def __init__(self, horizon: int = 8, context_dim: int = 0, backbone_dim: int = 512):
        super().__init__()
        self.horizon = horizon
        self.context_dim = context_dim if context_dim > 0 else 0
        self.backbone_dim = backbone_dim
        self.backbone = EasyBackbone(in_channels=3, out_dim=backbone_dim)
        head_in_dim = backbone_dim + self.context_dim
        hidden = 512
        self.hidden = hidden
        self.head_wp = nn.Sequential(
            nn.Linear(head_in_dim, hidden), nn.ReLU(inplace=True),
            nn.Linear(hidden, hidden), nn.ReLU(inplace=True),
            nn.Linear(hidden, horizon * 2),
        )
        self.head_spd = nn.Sequential(
            nn.Linear(head_in_dim, hidden), nn.ReLU(inplace=True),
            nn.Linear(hidden, hidden), nn.ReLU(inplace=True),
            nn.Linear(hidden, horizon),
        )
        self.precision = "auto"
        self._packs = {}

    def _pack(self, dtype, device):
        stamp = params_stamp([self])
        key = (dtype, device.index)
        p = self._packs.get(key)
        if p is None or p["stamp"] != stamp:
            net = self.backbone.net
            convs = [
                {},      # conv1 (Cin=3), packed on demand per stem mode (see _first)
                _ops.pack_conv([net[3]], [net[4]], dtype, device, relu=True),
                _ops.pack_conv([net[6]], [net[7]], dtype, device, relu=True),
                _ops.pack_conv([net[9]], [net[10]], dtype, device, relu=True),
            ]
            lin = [self.backbone.fc, self.head_wp[0], self.head_wp[2], self.head_wp[4],
                   self.head_spd[0], self.head_spd[2], self.head_spd[4]]
            flat = _ops.flat_params([t for l in lin for t in (l.weight, l.bias)], device)
            p = dict(stamp=stamp, convs=convs, flat=flat, flat16=flat.to(torch.bfloat16))
            self._packs[key] = p
        return p

    def _first(self, p, mode: str, dtype, device):
        first = p["convs"][0].get(mode)
        if first is None:
            net = self.backbone.net
            first = (_ops.pack_stem([net[0]], [net[1]], device, relu=True) if mode == "tc" else
                     _ops.pack_rowwin([net[0]], [net[1]], device, relu=True) if mode == "rowwin" else
                     _ops.pack_conv([net[0]], [net[1]], dtype, device, relu=True, cin_pad=4))
            p["convs"][0][mode] = first
        return first

    def backbone_features(self, image: torch.Tensor, _x_nhwc=None, _dtype=None, _conv1=None) -> torch.Tensor:
        """EasyBackbone.net (trajectory_head.py:8-21) on the inference kernels: the last convolution's NHWC map, or - bf16
        tensor-core mode - its spatial mean [B,1,1,256] fp32 (the pooling the head would run first).  Independent of the
        gate, so AutoMoE.forward can run it on a second stream beside the gating kernel."""
        dtype = _dtype or resolve_dtype(self.precision)
        p = self._pack(dtype, image.device)
        B, _, H, W = image.shape
        h, w = H, W
        if _conv1 is not None:       # conv1 already computed by the caller (fused into the experts' stem GEMM)
            x, convs = _conv1, p["convs"][1:]
            h, w = x.shape[1], x.shape[2]
        else:
            x = _x_nhwc if _x_nhwc is not None else stage_image(image, dtype)
            convs = [self._first(p, _ops.stem_mode(dtype, H, W), dtype, image.device)] + p["convs"][1:]
        for pc in convs:
            if isinstance(pc, _ops.PackedStem):
                x = _ops.stem_forward(pc, x, B, h, w)[0]
            elif isinstance(pc, _ops.PackedRowwin):
                x = _ops.conv2d_rowwin(pc, x, B, h, w)
            else:
                x = _ops.conv2d(pc, x, B, h, w)
            h, w = x.shape[1], x.shape[2]
        if _ops.mlp_tc(dtype) and B >= 16 and x.dtype == torch.bfloat16 and x.shape[3] % 8 == 0:
            x = _ops.mean_hw_nhwc(x).view(B, 1, 1, x.shape[3])
        return x

    def forward(self, image: torch.Tensor, context: Optional[torch.Tensor] = None, _x_nhwc=None,
                _dtype=None, _conv1=None, _feat=None) -> Dict[str, torch.Tensor]:
        from .._train_forward import policy_forward, wants_grad
        if wants_grad(self):
            return policy_forward(self, image, context)
        if not image.is_cuda:
            raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
        dtype = _dtype or resolve_dtype(self.precision)
        p = self._pack(dtype, image.device)
        x = _feat if _feat is not None else self.backbone_features(image, _x_nhwc, dtype, _conv1)
        if context is not None:
            if self.context_dim == 0:
                raise ValueError("TrajectoryPolicy was built with context_dim=0 but a context was given")
            cvec = context.float().contiguous()
            cdim = self.context_dim
        else:
            if self.context_dim != 0:
                raise ValueError("TrajectoryPolicy expects a context of dim %d" % self.context_dim)
            cvec, cdim = None, 0
        wp, spd = _ops.policy_head(x, cvec, p["flat"], self.backbone_dim, cdim, self.hidden, self.horizon,
                                   params_bf16=p["flat16"] if _ops.mlp_tc(dtype) else None)
        return {"waypoints": wp.view(-1, self.horizon, 2), "speed": spd.view(-1, self.horizon)}
