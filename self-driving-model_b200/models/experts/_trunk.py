"""ResNet-18 trunk + 2-conv head shared by the three BDD experts, run through the
sm_100a kernels for G experts at once ("grouped": activations stacked on the batch axis,
weights on the Cout axis).

Parameter holders reproduce the attribute names of torchvision.models.resnet18 sliced
with children()[:-2] (reference: models/experts/bdd_detection_expert.py:9-10), so the
state_dict keys are identical (experts.N.backbone.4.0.conv1.weight, ...).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from ... import _ops
from ..._cabi import check, ctx, lib
from ..._ops import ptr, stream_ptr


class ParamHolder(nn.Sequential):
    """Container that owns parameters but is never executed: the math lives in the kernels."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("parameter holder: call the owning expert/model, not this container")


class BasicBlock(nn.Module):
    """torchvision.models.resnet.BasicBlock attribute layout (conv1,bn1,relu,conv2,bn2,downsample)."""

    def __init__(self, inplanes: int, planes: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = ParamHolder(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("parameter holder: call the owning expert/model, not this block")


def make_resnet18_trunk(pretrained: bool = False) -> ParamHolder:
    """children()[:-2] of resnet18: conv1, bn1, relu, maxpool, layer1..layer4 (indices 0..7)."""
    layers = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
              nn.MaxPool2d(3, 2, 1)]
    inplanes = 64
    for planes, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
        layers.append(ParamHolder(BasicBlock(inplanes, planes, stride), BasicBlock(planes, planes, 1)))
        inplanes = planes
    trunk = ParamHolder(*layers)
    # torchvision init (resnet.py: kaiming_normal_ fan_out/relu for convs, BN weight 1 / bias 0)
    for m in trunk.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    if pretrained:
        try:
            import torchvision.models as tvm  # the reference's source of ImageNet weights
            sd = tvm.resnet18(pretrained=True).state_dict()
            names = ["conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4"]
            remap = {}
            for k, v in sd.items():
                head, _, rest = k.partition(".")
                if head in names:
                    remap[f"{names.index(head)}.{rest}"] = v
            trunk.load_state_dict(remap)
        except Exception as e:  # no network in most deployments; a checkpoint is loaded afterwards
            import warnings
            warnings.warn(f"pretrained_backbone=True but ImageNet weights are unavailable ({e}); using random init")
    return trunk


def make_head(out_channels: int) -> ParamHolder:
    """nn.Sequential(Conv2d(512,256,3,padding=1), ReLU, Conv2d(256,out,1)) — bdd_*_expert.py:12-16."""
    return ParamHolder(nn.Conv2d(512, 256, kernel_size=3, padding=1), nn.ReLU(), nn.Conv2d(256, out_channels, kernel_size=1))


@dataclass
class TrunkPack:
    dtype: torch.dtype
    G: int
    stems: dict             # stem mode ('tc' | 'rowwin' | 'simt') -> packed first-layer filters (filled on demand)
    stem_src: tuple         # (convs, bns, device) the stems are packed from
    blocks: list            # [(conv1, conv2, down|None)]
    head3: _ops.PackedConv  # 3x3 512->256 + bias + ReLU (grouped)
    head1_w: List[torch.Tensor]  # per expert [N,256] fp32
    head1_b: List[torch.Tensor]
    n_ch: List[int]
    stamp: tuple


def params_stamp(modules) -> tuple:
    """Cheap change detector for cached packs: in-place updates bump Tensor._version,
    .to()/load swaps change data_ptr."""
    v = 0
    first = None
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            v += t._version
            if first is None:
                first = t.data_ptr()
    return (v, first)


def pack_trunks(experts, heads, dtype: torch.dtype, device) -> TrunkPack:
    """experts: list of modules with .backbone (ParamHolder trunk); heads: their 2-conv heads, or None for a
    trunk-only pack (the nuScenes expert pools the trunk output instead)."""
    G = len(experts)
    bbs = [e.backbone for e in experts]
    blocks = []
    for li in range(4, 8):
        for bi in range(2):
            blks = [bb[li][bi] for bb in bbs]
            c1 = _ops.pack_conv([b.conv1 for b in blks], [b.bn1 for b in blks], dtype, device, relu=True)
            c2 = _ops.pack_conv([b.conv2 for b in blks], [b.bn2 for b in blks], dtype, device, relu=True)
            dn = None
            if blks[0].downsample is not None:
                dn = _ops.pack_conv([b.downsample[0] for b in blks], [b.downsample[1] for b in blks], dtype, device,
                                    relu=False)
            blocks.append((c1, c2, dn))
    if heads is None:
        return TrunkPack(dtype, G, {}, ([bb[0] for bb in bbs], [bb[1] for bb in bbs], device), blocks, None, [], [], [],
                         params_stamp(list(experts)))
    head3 = _ops.pack_conv([h[0] for h in heads], None, dtype, device, relu=True)
    w1 = [h[2].weight.detach().to(device=device, dtype=torch.float32).reshape(h[2].weight.shape[0], -1).contiguous()
          for h in heads]
    b1 = [h[2].bias.detach().to(device=device, dtype=torch.float32).contiguous() for h in heads]
    return TrunkPack(dtype, G, {}, ([bb[0] for bb in bbs], [bb[1] for bb in bbs], device), blocks, head3, w1, b1,
                     [w.shape[0] for w in w1], params_stamp(list(experts)))


def stem_pack(pack: TrunkPack, mode: str):
    """First-layer filters of all experts packed for `mode` (cached inside the TrunkPack)."""
    st = pack.stems.get(mode)
    if st is None:
        convs, bns, device = pack.stem_src
        if mode == "tc":       # Cin=3 stems of all experts as one GEMM over the raw image rows
            st = _ops.pack_stem(convs, bns, device, relu=True)
        elif mode == "rowwin":
            st = _ops.pack_rowwin(convs, bns, device, relu=True)
        else:
            st = _ops.pack_conv(convs, bns, pack.dtype, device, relu=True, cin_pad=4)
        pack.stems[mode] = st
    return st


def run_trunk_train(expert, image: torch.Tensor, with_head: bool = True) -> torch.Tensor:
    """Differentiable forward of ONE expert (ResNet-18 trunk + 2-conv head) for expert training
    (training/train_bdd100k_ddp.py:117-186): fp32 NHWC, every layer an autograd shim over
    the sm_100a training kernels (training/functional.py).  BatchNorm follows module.training (batch
    statistics + running-stat update in train mode).  Returns the head output [B,h,w,N] (NHWC).

    torchvision ResNet._forward_impl / BasicBlock.forward (resnet.py:89-105,266-278), bdd_*_expert.py:12-24."""
    from ...training import functional as TF
    if not image.is_cuda:
        raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
    bb = expert.backbone
    x = TF.image_nhwc4(image)
    y = TF.conv_bn_act(x, bb[0], bb[1], relu=True)
    y = TF.max_pool3x3s2(y)
    for li in range(4, 8):
        for blk in bb[li]:
            out = TF.conv_bn_act(y, blk.conv1, blk.bn1, relu=True)
            out = TF.conv_bn_act(out, blk.conv2, blk.bn2, relu=False)
            idn = y if blk.downsample is None else TF.conv_bn_act(y, blk.downsample[0], blk.downsample[1], relu=False)
            y = TF.add_relu(out, idn)
    if not with_head:
        return y                                                          # [B,h,w,512] trunk output
    head = expert.head_module()
    h = TF.conv_bn_act(y, head[0], None, relu=True)
    return TF.conv_bn_act(h, head[2], None, relu=False)


# ---- frozen experts in the reference's train mode, G experts in lockstep ---------------------------------------------------
_CONST_CACHE: dict = {}


def _const_vec(device: torch.device, n: int, value: float) -> torch.Tensor:
    """ones / zeros [n] fp32 (identity scale / zero bias of a plain convolution), one allocation per (device, n, value)."""
    key = (device.index, n, value)
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.full((n,), value, device=device, dtype=torch.float32)
        _CONST_CACHE[key] = t
    return t


def _grouped_split_weight(convs) -> torch.Tensor:
    """[G*Cout, KH, KW, 6, Cin] bf16: the six-term split packs (training/functional._split_weight) of G same-shape convolutions
    back to back - the weight layout of amoe_conv2d_fwd_f32tc_grouped.  Cached on the first weight until any of them changes."""
    from ...training import functional as TF
    ws = [c.weight for c in convs]
    key = tuple((w._version, w.data_ptr()) for w in ws)
    hit = getattr(ws[0], "_amoe_split6_grouped", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    out = torch.cat([TF._split_weight(w, False) for w in ws], dim=0).contiguous()
    try:
        ws[0]._amoe_split6_grouped = (key, out)
    except Exception:
        pass
    return out


def grouped_train_supported(experts, image: torch.Tensor) -> bool:
    """True when run_trunks_train_grouped takes these experts: at least two ResNet-18 experts of one geometry, frames whose
    stage inputs stay even (every convolution behind the stem is then a split-operand tensor-core launch)."""
    from ...training import functional as TF
    if len(experts) < 2 or len(experts) > 4 or not TF.train_tc() or image.dim() != 4:
        return False
    H, W = image.shape[2], image.shape[3]
    if H % 32 != 0 or W % 32 != 0:
        return False
    bb0 = experts[0].backbone
    bn0 = bb0[1]
    for e in experts:
        bb = e.backbone
        if any(p.requires_grad for p in e.parameters()):
            return False
        # every BatchNorm of the trunk on batch statistics, affine, tracked, one momentum / eps: what the grouped passes compute
        # (anything else - a layer left in eval mode, momentum=None - keeps the layer-by-layer path that reads each module's flags)
        for m in bb.modules():
            if isinstance(m, nn.BatchNorm2d) and not (m.training and m.affine and m.track_running_stats and m.momentum is not None
                                                      and m.momentum == bn0.momentum and m.eps == bn0.eps):
                return False
        for li in range(4, 8):
            if len(bb[li]) != len(bb0[li]):
                return False
            for blk, blk0 in zip(bb[li], bb0[li]):
                if blk.conv1.weight.shape != blk0.conv1.weight.shape or (blk.downsample is None) != (blk0.downsample is None):
                    return False
        if e.head_module()[0].weight.shape != experts[0].head_module()[0].weight.shape:
            return False
    return True


def _conv_grouped(x: torch.Tensor, convs, relu: bool = False) -> torch.Tensor:
    """G same-shape nn.Conv2d layers on x [G,B,H,W,Cin] fp32 -> [G,B,Ho,Wo,Cout] fp32: ONE split-operand tensor-core launch
    (csrc/conv_tc.cu, fp32-accurate); bias (+ReLU) in the epilogue when the layers have one."""
    from ...training import functional as TF
    G, B, H, W, Cin = x.shape
    c0 = convs[0]
    Cout, _, KH, KW = c0.weight.shape
    stride, pad = c0.stride[0], c0.padding[0]
    if not lib().amoe_conv2d_f32tc_supported(H, W, Cin, Cout, KH, KW, stride):
        raise NotImplementedError(f"grouped training convolution {Cin}->{Cout} {KH}x{KW}/s{stride} on {H}x{W}")
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    dev = x.device
    xs, wsplit = TF._split3(x), _grouped_split_weight(convs)          # named: they outlive the launch
    if c0.bias is not None:
        bias = torch.cat([c.bias.detach().float() for c in convs]).contiguous()
    else:
        bias = _const_vec(dev, G * Cout, 0.0)
    y = torch.empty((G, B, Ho, Wo, Cout), device=dev, dtype=torch.float32)
    check(lib().amoe_conv2d_fwd_f32tc_grouped(ctx(dev), ptr(xs), ptr(wsplit), ptr(_const_vec(dev, G * Cout, 1.0)), ptr(bias), None,
                                              ptr(y), G, 0, B, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, int(relu),
                                              stream_ptr(dev)), "conv2d_fwd_f32tc_grouped")
    del xs
    return y


def _bn_train_grouped(x: torch.Tensor, bns, relu: bool) -> torch.Tensor:
    """G train-mode nn.BatchNorm2d layers on x [G,B,H,W,C]: batch statistics, running statistics of every layer updated in
    place (momentum of the layer), one launch per pass for all G (amoe_bn_train_fwd_grouped)."""
    import ctypes as C_
    G, B, H, W, C = x.shape
    M = B * H * W
    dev = x.device
    b0 = bns[0]
    if any(bn.momentum is None or bn.momentum != b0.momentum or bn.eps != b0.eps or not bn.track_running_stats or
           bn.weight is None for bn in bns):
        raise NotImplementedError("grouped BatchNorm needs affine layers with one momentum / eps and tracked running statistics")

    def arr(ts):
        for t in ts:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise RuntimeError("grouped BatchNorm takes contiguous fp32 parameters / buffers on the activation's device")
        return (C_.c_void_p * G)(*[t.data_ptr() for t in ts])
    y = torch.empty_like(x)
    ws = torch.empty(G * max(1, lib().amoe_colreduce_workspace_floats(M, C)), device=dev, dtype=torch.float32)
    mean = torch.empty(G * C, device=dev, dtype=torch.float32)
    rstd = torch.empty(G * C, device=dev, dtype=torch.float32)
    check(lib().amoe_bn_train_fwd_grouped(ctx(dev), ptr(x), arr([bn.weight.detach() for bn in bns]), arr([bn.bias.detach() for bn in bns]),
                                          arr([bn.running_mean for bn in bns]), arr([bn.running_var for bn in bns]),
                                          float(b0.momentum), float(b0.eps), ptr(y), ptr(mean), ptr(rstd), ptr(ws), G, M, C, int(relu),
                                          stream_ptr(dev)), "bn_train_fwd_grouped")
    for bn in bns:      # the kernel wrote the running statistics through raw pointers
        torch.autograd.graph.increment_version([bn.running_mean, bn.running_var])
    return y


def run_trunks_train_grouped(experts, image: torch.Tensor):
    """Frozen ResNet-18 experts exactly as the reference runs them inside model.train() (training/train_gating_network.py:85:
    batch-statistics BatchNorm, running statistics and num_batches_tracked updated) - all G experts layer by layer in
    LOCKSTEP: behind the per-expert first layers every convolution, BatchNorm pass and residual add is one launch over
    [G,B,H,W,C].  At 32 frames per GPU a layer3 / layer4 convolution of ONE expert has 64 / 32 output tiles for 148 SMs;
    grouped, three experts cost little more than one.  No autograd graph (the experts are frozen); same kernels and arithmetic
    as run_trunk_train, expert by expert.  Returns the head outputs [B,h,w,N_g] (NHWC fp32), one per expert.

    torchvision ResNet._forward_impl / BasicBlock.forward (resnet.py:89-105,266-278), bdd_*_expert.py:12-24."""
    from ...training import functional as TF
    G = len(experts)
    bbs = [e.backbone for e in experts]
    tracked = []
    with torch.no_grad():
        x = TF.image_nhwc4(image)
        pooled = [TF.max_pool3x3s2(TF.conv_bn_act(x, bb[0], bb[1], relu=True)) for bb in bbs]     # first layers: per expert
        y = torch.stack(pooled, dim=0)                                                           # [G,B,H/4,W/4,64]
        del pooled
        for li in range(4, 8):
            for bi in range(len(bbs[0][li])):
                blks = [bb[li][bi] for bb in bbs]
                out = _bn_train_grouped(_conv_grouped(y, [b.conv1 for b in blks]), [b.bn1 for b in blks], relu=True)
                out = _bn_train_grouped(_conv_grouped(out, [b.conv2 for b in blks]), [b.bn2 for b in blks], relu=False)
                tracked += [b.bn1 for b in blks] + [b.bn2 for b in blks]
                if blks[0].downsample is not None:
                    idn = _bn_train_grouped(_conv_grouped(y, [b.downsample[0] for b in blks]), [b.downsample[1] for b in blks],
                                            relu=False)
                    tracked += [b.downsample[1] for b in blks]
                else:
                    idn = y
                y = TF.add_relu(out, idn)
        torch._foreach_add_([bn.num_batches_tracked for bn in tracked], 1)       # what 3 x 19 BatchNorm forwards do one by one
        heads = [e.head_module() for e in experts]
        h = _conv_grouped(y, [hd[0] for hd in heads], relu=True)                 # 3x3 512 -> 256 + bias + ReLU
        return [TF.conv_bn_act(h[g], heads[g][2], None, relu=False) for g in range(G)]   # 1x1 heads: one channel count each


def _stem_f32_tc(pack: TrunkPack, stem, x_nhwc: torch.Tensor, B: int, H: int, W: int) -> Optional[torch.Tensor]:
    """fp32 (parity) mode: the Cin = 3 first layers of all experts fp32-accurate on the tensor cores - the Toeplitz GEMM with
    three-way split operands the training forward uses (stem_tc_f32_kernel, csrc/stem_tc.cu; 2.5-3.2e-7 against fp64), with the
    folded BatchNorm scale / bias and the ReLU in its epilogue.  One launch per expert on one shared split frame.  Returns
    [G*B,H/2,W/2,Cout] fp32, or None for shapes the kernel does not take (the CUDA-core convolution then runs: 1.6 ms for three
    experts at 32 frames against 3 x 0.1 ms)."""
    from ...training import functional as TF
    convs = pack.stem_src[0]
    c0 = convs[0]
    Cout, Cin, KH, KW = c0.weight.shape
    if pack.dtype != torch.float32 or not _ops.f32_tc() or x_nhwc.dtype != torch.float32 or x_nhwc.dim() != 4 or x_nhwc.shape[-1] != 4:
        return None
    if c0.stride[0] != c0.stride[1] or c0.padding[0] != c0.padding[1] or \
            not TF._stem_tc_ok(4, Cin, Cout, KH, KW, c0.stride[0], c0.padding[0], H, W):
        return None
    dev = x_nhwc.device
    Ho, Wo = (H + 2 * c0.padding[0] - KH) // 2 + 1, (W + 2 * c0.padding[0] - KW) // 2 + 1
    xs = TF._stem_split_frame(x_nhwc.contiguous())                       # named: read by every launch below
    y = torch.empty((len(convs) * B, Ho, Wo, Cout), device=dev, dtype=torch.float32)
    for g, conv in enumerate(convs):
        ws = TF._stem_split_weight(conv.weight, conv.padding[0])
        check(lib().amoe_stem_fwd_f32tc(ctx(dev), ptr(xs), ptr(ws), ptr(stem.scale[g * Cout:]), ptr(stem.bias[g * Cout:]),
                                        ptr(y[g * B:]), B, H, W, xs.shape[2], _ops.STEM_KH, Cout, int(stem.relu), stream_ptr(dev)),
              "stem_fwd_f32tc")
    return y


def stage_image(image: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """NCHW fp32 frame -> the NHWC layout the first convolutions read (see _ops.stem_mode): the
    physically padded bf16 frame of the tensor-core stem, padded rows for the row-window variant, or
    plain [B,H,W,4] for the CUDA-core kernel."""
    H, W = image.shape[2], image.shape[3]
    mode = _ops.stem_mode(dtype, H, W)
    if mode == "tc":
        if not _ops.stem_supported(H, W):
            raise NotImplementedError(f"AMOE_STEM=tc needs even H, W and W <= 256 (got {H}x{W})")
        return _ops.stage_image_stem(image)
    if mode == "rowwin":
        return _ops.image_to_nhwc_padded(image, _ops.ROWWIN_CP, _ops.ROWWIN_LEFT, _ops.rowwin_wpad(W), dtype)
    return _ops.image_to_nhwc(image, 4, dtype)


def trunk_pool_pad(pack: TrunkPack, H: int, W: int) -> int:
    """1 if the max-pooled stem output must be stored with a zero border (its consumers are the
    halo-reuse 3x3 kernels), else 0."""
    flat_ok = pack.dtype == torch.bfloat16 and _ops.use_flat()
    h2, w2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
    return 1 if (flat_ok and _ops.flat_supported(pack.blocks[0][1], h4, w4, pack.dtype)) else 0


def chunked_stem_layer1_supported(pack: TrunkPack, B: int, H: int, W: int) -> bool:
    """True when stem+max-pool and the four layer1 convolutions can walk the batch in L2-sized chunks."""
    chunk = _ops.l2_chunk_images()
    if chunk <= 0 or B <= chunk or pack.dtype != torch.bfloat16 or not _ops.use_flat():
        return False
    if _ops.stem_mode(pack.dtype, H, W) != "tc" or not _ops.stem_pool_supported(H, W) or not trunk_pool_pad(pack, H, W):
        return False
    (c1a, c2a, dna), (c1b, c2b, dnb) = pack.blocks[0], pack.blocks[1]
    if dna is not None or dnb is not None or c1a.sh != 1 or c1b.sh != 1:
        return False
    return all(_ops.flat_supported(c, H // 4, W // 4, pack.dtype) for c in (c1a, c2a, c1b, c2b))


def run_stem_layer1_chunked(pack: TrunkPack, fused_stem, x_nhwc: torch.Tensor, B: int, H: int, W: int,
                            rest_out: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """conv1+bn1+relu+maxpool and layer1 (two BasicBlocks) of all experts, one sub-batch at a time.

    Every layer1 convolution moves ~3.3 MB per frame (read + write, 3 experts) for 0.9 GFLOP, so at full
    batch the stage runs at HBM speed.  Walking the batch in chunks whose activations fit the 126 MB L2
    lets each convolution read what the previous launch just wrote from L2; only the stage's final
    output goes to HBM, written straight into the full [G*B,...] tensor (strided expert groups).
    fused_stem: PackedStem whose first G convolutions are the expert stems (more may follow: their
    full-resolution outputs go to rest_out, sliced per chunk)."""
    G = pack.G
    chunk = _ops.l2_chunk_images()
    h4, w4 = H // 4, W // 4
    (c1a, c2a, _), (c1b, c2b, _) = pack.blocks[0], pack.blocks[1]
    dev = x_nhwc.device
    y_full = torch.empty((G * B, h4 + 2, w4 + 2, c2b.cout), device=dev, dtype=torch.bfloat16)
    bufs = {}
    for b0 in range(0, B, chunk):
        bc = min(chunk, B - b0)
        if bc not in bufs:   # the last chunk may be shorter
            bufs[bc] = [torch.empty((G * bc, h4 + 2, w4 + 2, 64), device=dev, dtype=torch.bfloat16) for _ in range(4)]
        p0, t1, t2, t3 = bufs[bc]
        rest = [r[b0:b0 + bc] for r in rest_out] if rest_out is not None else None
        _ops.stem_pool_forward(fused_stem, x_nhwc[b0:b0 + bc], bc, H, W, G, 1, pooled=p0, rest=rest)
        _ops.conv3x3_flat(c1a, p0, bc, h4, w4, out=t1)
        _ops.conv3x3_flat(c2a, t1, bc, h4, w4, residual=p0, out=t2)
        _ops.conv3x3_flat(c1b, t2, bc, h4, w4, out=t3)
        _ops.conv3x3_flat(c2b, t3, bc, h4, w4, residual=t2, out=y_full[b0:], out_group_images=B)
    return y_full


def run_trunks(pack: TrunkPack, image: torch.Tensor, x_nhwc: Optional[torch.Tensor] = None,
               stem_out: Optional[torch.Tensor] = None, stem_pooled: Optional[torch.Tensor] = None,
               layer1_out: Optional[torch.Tensor] = None, features_only: bool = False, after_head3=None):
    """image: [B,3,H,W] fp32 NCHW.  Returns (low_res list of [B,h,w,N_e] fp32, pooled [B,sumC] fp32, (h,w)).
    after_head3: optional callable invoked once the last tensor-bound launch (the heads' 3x3 convolution) is enqueued -
    the point where AutoMoE.forward forks the policy backbone onto its own stream.

    Follows torchvision ResNet._forward_impl up to layer4 and BasicBlock.forward
    (conv-bn-relu-conv-bn + identity/downsample, add, relu), then the expert head
    (bdd_detection_expert.py:18-20).
    """
    B, _, H, W = image.shape
    G = pack.G
    pad = trunk_pool_pad(pack, H, W)
    stem = stem_pack(pack, _ops.stem_mode(pack.dtype, H, W))
    if layer1_out is None and stem_pooled is None and stem_out is None and isinstance(stem, _ops.PackedStem) \
            and chunked_stem_layer1_supported(pack, B, H, W):
        if x_nhwc is None:
            x_nhwc = stage_image(image, pack.dtype)
        layer1_out = run_stem_layer1_chunked(pack, stem, x_nhwc, B, H, W)
    if layer1_out is not None or stem_pooled is not None:
        y = None                                                         # stem + max-pool (+ layer1) already done
    elif stem_out is not None:
        y = stem_out                                                     # computed by the caller (fused with policy conv1)
    else:
        if x_nhwc is None:
            x_nhwc = stage_image(image, pack.dtype)
        if isinstance(stem, _ops.PackedStem) and _ops.stem_pool_supported(H, W):
            stem_pooled = _ops.stem_pool_forward(stem, x_nhwc, B, H, W, G, pad)[0]
            y = None
        elif isinstance(stem, _ops.PackedStem):
            y = _ops.stem_forward(stem, x_nhwc, B, H, W, groups=[G])[0]   # [G*B,H/2,W/2,64]
        elif isinstance(stem, _ops.PackedRowwin):
            y = _ops.conv2d_rowwin(stem, x_nhwc, B, H, W)
        else:
            y = _stem_f32_tc(pack, stem, x_nhwc, B, H, W)                  # fp32 mode: split-operand Toeplitz GEMM when it applies
            if y is None:
                y = _ops.conv2d(stem, x_nhwc, B, H, W, x_shared=True)
    # Activations of the 64/128-channel stages live in a physically padded layout (zero border of one
    # pixel) when their 3x3/s1 convolutions run through the halo-reuse kernel; `pad` tracks the layout.
    flat_ok = pack.dtype == torch.bfloat16 and _ops.use_flat()
    h2, w2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    h_cur, w_cur = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
    blocks = pack.blocks
    if layer1_out is not None:
        y, blocks, pad = layer1_out, pack.blocks[2:], 1                  # [G*B,H/4+2,W/4+2,64], layer1 done
    elif stem_pooled is not None:
        y = stem_pooled                                                  # [G*B,H/4(+2),W/4(+2),64]
    else:
        y = _ops.maxpool3x3s2(y, out_pad=pad)
    for (c1, c2, dn) in blocks:
        # geometry of this block's output and whether its stride-1 convs take the flat kernel
        h_out = (h_cur + 2 * c1.ph - c1.kh) // c1.sh + 1
        w_out = (w_cur + 2 * c1.pw - c1.kw) // c1.sw + 1
        pad_out = 1 if (flat_ok and _ops.flat_supported(c2, h_out, w_out, pack.dtype)) else 0
        if pad and pad_out and c1.sh == 1 and _ops.flat_supported(c1, h_cur, w_cur, pack.dtype):
            out = _ops.conv3x3_flat(c1, y, B, h_cur, w_cur)
            identity = y if dn is None else _ops.conv2d(dn, y, B, h_cur, w_cur, in_pad=pad, out_pad=pad_out)
        elif _ops.dual_supported(c1, dn, h_cur, w_cur, pad, pack.dtype):
            out, identity = _ops.conv2d_dual(c1, dn, y, B, h_cur, w_cur, in_pad=pad, out_pad=pad_out)   # stage entry: one launch
        else:
            out = _ops.conv2d(c1, y, B, h_cur, w_cur, in_pad=pad, out_pad=pad_out, zero_border=True)
            identity = y if dn is None else _ops.conv2d(dn, y, B, h_cur, w_cur, in_pad=pad, out_pad=pad_out)
        if pad_out:
            y = _ops.conv3x3_flat(c2, out, B, h_out, w_out, residual=identity)
        else:
            y = _ops.conv2d(c2, out, B, h_out, w_out, residual=identity)
        h_cur, w_cur, pad = h_out, w_out, pad_out
    h, w = y.shape[1], y.shape[2]
    if features_only:
        return y                                                         # [G*B,h,w,512] (layer4 is never stored padded)
    hid = _ops.conv2d(pack.head3, y, B, h, w)                            # [G*B,h,w,256]
    if after_head3 is not None:
        after_head3()
    pooled = torch.empty((B, sum(pack.n_ch)), device=image.device, dtype=torch.float32)
    lows, off = [], 0
    for g in range(G):
        lows.append(_ops.head1x1_pool(hid[g * B:(g + 1) * B], pack.head1_w[g], pack.head1_b[g], pooled, off))
        off += pack.n_ch[g]
    return lows, pooled, (h, w)
