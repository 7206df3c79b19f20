"""Common forward of the three BDD perception experts (grouped or stand-alone)."""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from ... import _ops
from .._precision import resolve_dtype
from ._trunk import TrunkPack, pack_trunks, params_stamp, run_trunks


class BDDExpertBase(nn.Module):
    """ResNet-18 trunk + 3x3/ReLU/1x1 head.  Sub-classes name the head attribute
    (`head` / `decoder`) exactly as the reference does so state_dict keys match."""

    head_attr = "head"
    upsample_to_input = False

    def __init__(self):
        super().__init__()
        self._packs = {}
        self.precision = "auto"

    def head_module(self):
        return getattr(self, self.head_attr)

    def out_channels(self) -> int:
        return self.head_module()[2].weight.shape[0]

    # -- reference forward signature: forward(x: [B,3,H,W]) --
    def forward(self, x: torch.Tensor):
        from .._train_forward import wants_grad
        if wants_grad(self):
            return self.forward_train(x)
        outs, _ = run_experts([self], x, resolve_dtype(self.precision), self._packs)
        return outs[0]

    def forward_train(self, x: torch.Tensor):
        """Differentiable forward (expert training): see _trunk.run_trunk_train."""
        from ._trunk import run_trunk_train
        low = run_trunk_train(self, x)                      # [B,h,w,N] NHWC fp32
        return self.format_output_train(low, x.shape[2], x.shape[3])

    def format_output_train(self, low: torch.Tensor, H: int, W: int):
        """Default (segmentation / drivable): differentiable x32 bilinear up-sampling to NCHW logits."""
        from ...training import functional as TF
        return TF.upsample_bilinear_nchw(low, H, W)

    def format_output(self, low: torch.Tensor, H: int, W: int, dtype: torch.dtype):
        raise NotImplementedError


def _check_eval(experts):
    for e in experts:
        if e.training:
            raise NotImplementedError(
                "automoe_b200 experts run eval-mode BatchNorm folded into the conv epilogue; "
                "train-mode (batch-statistics) BatchNorm is not implemented yet - call .eval() on the experts")


def get_trunk_pack(experts, dtype, device, cache: dict, with_head: bool = True) -> TrunkPack:
    stamp = params_stamp(experts)
    key = (dtype, device.index)
    pack: TrunkPack = cache.get(key)
    if pack is None or pack.stamp != stamp:
        pack = pack_trunks(experts, [e.head_module() for e in experts] if with_head else None, dtype, device)
        cache[key] = pack
    return pack


_SIDE_STREAMS: dict = {}


def side_stream(device: torch.device, which: int = 0) -> torch.cuda.Stream:
    """Auxiliary streams per device for work that overlaps the gate/policy tail of the forward
    (0: full-resolution logit writers, 1: policy backbone)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), which)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[key] = s
    return s


def _tensors_of(out):
    if isinstance(out, dict):
        return [t for t in out.values() if torch.is_tensor(t)]
    return [out] if torch.is_tensor(out) else []


def run_experts(experts: List[BDDExpertBase], image: torch.Tensor, dtype: torch.dtype, cache: dict, x_nhwc=None,
                stem_out=None, stem_pooled=None, layer1_out=None, overlap_outputs: bool = False,
                frozen_eval: bool = False, after_head3=None):
    """Run G experts on the same image batch in grouped launches.

    Returns (expert_outputs in the reference's format, dict(pooled=[B,sumC], n_ch=[...], exact_pool=bool)).
    models/automoe.py:156-187 (_run_experts) + the pooled statistics the extractors need.

    overlap_outputs: the x32 bilinear writers of the segmentation/drivable logits (HBM-bound, nothing
    downstream inside the forward reads them) are forked onto a side stream so the latency-bound gate and
    policy-head kernels run beside them; the caller must join with aux['join'] (main.wait_stream) before
    returning the outputs.
    """
    if not image.is_cuda:
        raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
    if not frozen_eval:      # frozen_eval: caller runs frozen experts inside a training step (eval BatchNorm)
        _check_eval(experts)
    pack = get_trunk_pack(experts, dtype, image.device, cache)
    B, _, H, W = image.shape
    lows, pooled, (h, w) = run_trunks(pack, image, x_nhwc, stem_out, stem_pooled, layer1_out, after_head3=after_head3)
    exact_pool = all((not e.upsample_to_input) or (H % h == 0 and W % w == 0) for e in experts)
    join = None
    if overlap_outputs and exact_pool:
        main = torch.cuda.current_stream(image.device)
        side = side_stream(image.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            outs = [e.format_output(low, H, W, dtype) for e, low in zip(experts, lows)]
        for low in lows:
            low.record_stream(side)
        for out in outs:
            for t in _tensors_of(out):
                t.record_stream(main)
        join = side
    else:
        outs = [e.format_output(low, H, W, dtype) for e, low in zip(experts, lows)]
    # mean over the up-sampled map == mean over the low-res map only for integer scale factors
    off = 0
    for e, out, n in zip(experts, outs, pack.n_ch):
        if e.upsample_to_input and (H % h != 0 or W % w != 0):
            pooled[:, off:off + n] = _ops.mean_hw_nchw(out)
        off += n
    return outs, dict(pooled=pooled, n_ch=pack.n_ch, join=join)
