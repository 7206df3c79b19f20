"""NuScenesExpert (image-only) — same constructor arguments, attribute names, state_dict keys and outputs as
models/experts/nuscenes_expert.py:96-190; the forward runs on the sm_100a kernels.

    image -> ResNet-18 (children()[:-1]: trunk + global average pool) -> Linear(512,256)
          -> + query_embed[q] -> Linear(256,256)+ReLU+Dropout -> Linear(256,128)+ReLU+Dropout      (per query)
          -> class_head Linear(128,10), bbox_head Linear(128,bbox_dim)

The trunk reuses the tcgen05 implicit-GEMM convolutions of the BDD experts (a group of one), the pooled feature is
one reduction kernel, and the multi-query head is three GEMM launches: the first decoder layer is linear in
(feature + query), so it splits into a per-frame product W h0[b] and a per-query constant W E[q] + bias that are
added (and rectified) by one broadcast kernel instead of multiplying B*Q rows by a 256x256 matrix.

The LiDAR branch (PointNet / TNet, use_lidar=True) is not on the path the shipped configuration takes
(models/configs/automoe/model_config.json: use_lidar false) and is rejected explicitly.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from ... import _ops
from .._precision import resolve_dtype
from ._trunk import ParamHolder, make_resnet18_trunk, params_stamp


class NuScenesExpert(nn.Module):
    def __init__(self, image_backbone=None, lidar_backbone=None, fusion: str = 'concat', num_queries: int = 100,
                 use_lidar: bool = False, use_tnet: bool = False, bbox_dim: int = 7, pretrained_backbone: bool = True):
        super().__init__()
        if image_backbone is not None or lidar_backbone is not None:
            raise NotImplementedError("NuScenesExpert: custom image/lidar backbones are not supported by the sm_100a path")
        if use_lidar:
            raise NotImplementedError("NuScenesExpert(use_lidar=True): the PointNet branch is outside the B200 hot path "
                                      "(the shipped model_config.json sets use_lidar=false)")
        trunk = make_resnet18_trunk(pretrained_backbone)
        # children()[:-1] of torchvision resnet18: indices 0..7 as the BDD trunks + the average pool at index 8
        self.image_backbone = ParamHolder(*list(trunk.children()), nn.AdaptiveAvgPool2d((1, 1)))
        self.image_projection = nn.Linear(512, 256)
        self.use_lidar = False
        self.lidar_backbone = None
        self.fusion_type = fusion
        fusion_dim = 256
        self.num_queries = num_queries
        self.bbox_dim = bbox_dim
        self.query_embed = nn.Embedding(num_queries, fusion_dim)
        self.decoder = nn.Sequential(
            nn.Linear(fusion_dim, 256), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(256, 128), nn.ReLU(), nn.Dropout(0.3),
        )
        self.class_head = nn.Linear(128, 10)
        self.bbox_head = nn.Linear(128, self.bbox_dim)
        self.precision = "auto"
        self._packs = {}
        self._head_cache = None

    # the trunk machinery of the BDD experts addresses the ResNet through `.backbone`
    @property
    def backbone(self):
        return self.image_backbone

    def _head_constants(self, device):
        """Per-query constant of the first decoder layer, W E[q] + bias, and the concatenated output heads."""
        stamp = (params_stamp([self.query_embed, self.decoder, self.class_head, self.bbox_head]), device)
        if self._head_cache is None or self._head_cache[0] != stamp:
            from ...training import functional as TF
            with torch.no_grad():
                c = TF._Linear.apply(self.query_embed.weight, self.decoder[0].weight, self.decoder[0].bias, False, 0.0, 0)
                w_out = torch.cat([self.class_head.weight, self.bbox_head.weight], dim=0).contiguous()
                b_out = torch.cat([self.class_head.bias, self.bbox_head.bias], dim=0).contiguous()
            self._head_cache = (stamp, c, w_out, b_out)
        return self._head_cache[1:]

    def head_forward(self, feat: torch.Tensor, train_dropout: bool = False) -> Dict[str, torch.Tensor]:
        """feat: [B,512] fp32 pooled trunk feature -> {'class_logits': [B,Q,10], 'bbox_preds': [B,Q,bbox_dim]}."""
        from ...training import functional as TF
        B, Q, dev = feat.shape[0], self.num_queries, feat.device
        p = float(self.decoder[2].p) if train_dropout else 0.0
        lin = TF._Linear.apply
        h0 = lin(feat, self.image_projection.weight, self.image_projection.bias, False, 0.0, 0)          # [B,256]
        if p == 0.0:
            c, w_out, b_out = self._head_constants(dev)
            a = lin(h0, self.decoder[0].weight, None, False, 0.0, 0)                                     # [B,256]
            h1 = torch.empty((B, Q, 256), device=dev, dtype=torch.float32)
            _ops.check(_ops.lib().amoe_bcast_add_relu(_ops.ctx(dev), _ops.ptr(a), _ops.ptr(c), _ops.ptr(h1), B, Q, 256,
                                                      _ops.stream_ptr(dev)), "bcast_add_relu")
            h2 = lin(h1.view(B * Q, 256), self.decoder[3].weight, self.decoder[3].bias, True, 0.0, 0)   # [B*Q,128]
        else:
            # train mode of the reference: Dropout(0.3) behind both decoder layers - the layers run as written
            _, w_out, b_out = self._head_constants(dev)
            v = (h0.unsqueeze(1) + self.query_embed.weight.unsqueeze(0)).reshape(B * Q, 256)             # plumbing: broadcast add
            h1 = lin(v, self.decoder[0].weight, self.decoder[0].bias, True, p, TF.next_seed())
            h2 = lin(h1, self.decoder[3].weight, self.decoder[3].bias, True, p, TF.next_seed())
        out = lin(h2, w_out, b_out, False, 0.0, 0).view(B, Q, 10 + self.bbox_dim)
        return {'class_logits': out[..., :10], 'bbox_preds': out[..., 10:], '_flat': out}

    def features_from_image(self, image: torch.Tensor, dtype: torch.dtype, x_nhwc=None) -> torch.Tensor:
        """ResNet-18 trunk (inference kernels, BatchNorm running statistics) + global average pool -> [B,512] fp32."""
        from ._base import get_trunk_pack
        from ._trunk import run_trunks
        pack = get_trunk_pack([self], dtype, image.device, self._packs, with_head=False)
        y = run_trunks(pack, image, x_nhwc=x_nhwc, features_only=True)                                   # [B,h,w,512]
        B, h, w, Cc = y.shape
        if y.dtype == torch.bfloat16:
            return _ops.mean_hw_nhwc(y)
        out = torch.empty((B, Cc), device=y.device, dtype=torch.float32)
        _ops.check(_ops.lib().amoe_gap_fwd(_ops.ctx(y.device), _ops.ptr(y), _ops.ptr(out), B, h * w, Cc, _ops.stream_ptr(y.device)),
                   "gap_fwd")
        return out

    def forward(self, batch, _x_nhwc=None, _dtype=None):
        image = batch['image'] if isinstance(batch, dict) else batch
        if not image.is_cuda:
            raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
        if any(p.requires_grad for p in self.parameters()) and torch.is_grad_enabled():
            raise NotImplementedError("training the nuScenes expert is not implemented on the sm_100a path: freeze it "
                                      "(AutoMoE.freeze_experts()) or run it under torch.no_grad()")
        dtype = _dtype or resolve_dtype(self.precision)
        with torch.no_grad():
            if self.training:
                # reference semantics inside model.train(): batch-statistics BatchNorm (running statistics updated) and
                # active Dropout(0.3) in the decoder
                from ...training import functional as TF
                from ._trunk import run_trunk_train
                y = run_trunk_train(self, image, with_head=False)
                out = self.head_forward(TF.global_avg_pool(y), train_dropout=True)
            else:
                out = self.head_forward(self.features_from_image(image, dtype, _x_nhwc))
        return out
