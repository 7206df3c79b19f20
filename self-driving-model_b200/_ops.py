"""Tensor-level wrappers over the C-ABI (device pointers + current stream).

Everything here launches sm_100a kernels from libautomoe_b200.so; torch is used only
to own memory.  Activations are NHWC, stacked over experts on the batch axis.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _cabi
from ._cabi import check, ctx, dtype_code, lib, ptr, stream_ptr


# bench.py sets PROFILE to a list to get (kernel name, flops, start_event, end_event) per conv launch
PROFILE = None


def _al8(n: int) -> int:
    return (n + 7) & ~7


def flat_params(tensors: List[torch.Tensor], device) -> torch.Tensor:
    """Concatenate fp32 tensors, each padded to a multiple of 8 floats (tensors stay 16-byte aligned in a bf16 copy)."""
    parts = []
    for t in tensors:
        f = t.detach().to(device=device, dtype=torch.float32).reshape(-1)
        pad = _al8(f.numel()) - f.numel()
        if pad:
            f = torch.cat([f, f.new_zeros(pad)])
        parts.append(f)
    return torch.cat(parts).contiguous()


def image_to_nhwc(img: torch.Tensor, cp: int, dtype: torch.dtype) -> torch.Tensor:
    """[B,C,H,W] fp32 NCHW -> [B,H,W,cp] (zero padded channels)."""
    if img.dtype != torch.float32:
        img = img.float()
    img = img.contiguous()
    B, Cc, H, W = img.shape
    out = torch.empty((B, H, W, cp), device=img.device, dtype=dtype)
    check(lib().amoe_image_nchw_to_nhwc(ctx(img.device), ptr(img), ptr(out), B, Cc, H, W, cp,
                                        dtype_code(dtype), stream_ptr(img.device)), "image_nchw_to_nhwc")
    return out


def image_to_nhwc_padded(img: torch.Tensor, cp: int, left: int, wpad: int, dtype: torch.dtype, top: int = 0,
                         hpad: Optional[int] = None, pad_value: float = 0.0) -> torch.Tensor:
    """[B,C,H,W] fp32 NCHW -> [B,hpad,wpad,cp]: image pixel (h,w) at (top+h, left+w), zeros elsewhere;
    the padding channels (c >= C) hold pad_value at every position."""
    if img.dtype != torch.float32:
        img = img.float()
    img = img.contiguous()
    B, Cc, H, W = img.shape
    hpad = H if hpad is None else hpad
    out = torch.empty((B, hpad, wpad, cp), device=img.device, dtype=dtype)
    check(lib().amoe_image_nchw_to_nhwc_padded_v(ctx(img.device), ptr(img), ptr(out), B, Cc, H, W, cp, left, wpad, top, hpad,
                                                 dtype_code(dtype), float(pad_value), stream_ptr(img.device)),
          "image_nchw_to_nhwc_padded")
    return out


# ---- first-layer GEMM over raw image rows (csrc/stem_tc.cu) ----
STEM_LEFT, STEM_TOP, STEM_KH = 4, 3, 7


def stem_supported(H: int, W: int) -> bool:
    return H % 2 == 0 and W % 2 == 0 and W // 2 <= 128


def stem_wpad(W: int) -> int:
    return (W + 6 + 7) & ~7


def stage_image_stem(img: torch.Tensor) -> torch.Tensor:
    """Frame in the layout amoe_stem_fwd reads: [B,H+6,Wpad,4] bf16, 3 zero rows top/bottom, 4 zero px left.
    The 4th (padding) channel is 1.0 everywhere: the folded stem filters add their bias through it; the
    unfolded filters have zeros there, so it is inert for them."""
    H, W = img.shape[2], img.shape[3]
    return image_to_nhwc_padded(img, 4, STEM_LEFT, stem_wpad(W), torch.bfloat16, top=STEM_TOP, hpad=H + 6, pad_value=1.0)


def stem_fold() -> bool:
    """Fused stem+pool kernel with BatchNorm folded into the filters / bias on the padding channel
    (AMOE_STEM_FOLD=0: scale/bias applied in fp32 in the epilogue, weights rounded exactly like autocast's)."""
    import os
    return os.environ.get("AMOE_STEM_FOLD", "1") != "0"


@dataclass
class PackedStem:
    """First-layer filters of several convolutions sharing the frame, concatenated on N, in the
    shared-memory image [KH*4][n_total][8] bf16 of stem_tc.cu."""
    w: torch.Tensor
    scale: torch.Tensor
    bias: torch.Tensor
    couts: List[int]       # output channels of each convolution (64 per expert stem, 32 for policy conv1)
    n_total: int
    relu: bool
    true_macs_per_px: int  # sum over convs of cout*cin*kh*kw (algorithmic work per output pixel)
    w_folded: Optional[torch.Tensor] = None   # same image with scale folded in and the bias on channel 3 (see stem_tc.cu)


def pack_stem(convs, bns, device, relu: bool = True) -> PackedStem:
    """convs: stride-2 nn.Conv2d with Cin<=4 and kernel 7x7/p3 or 5x5/p2 (same output size)."""
    st, h = stream_ptr(device), ctx(device)
    couts = [int(c.weight.shape[0]) for c in convs]
    n_total = sum(couts)
    assert n_total % 32 == 0 and n_total <= 256 and all(co % 32 == 0 for co in couts)
    wk = torch.zeros((n_total, STEM_KH, 8, 4), dtype=torch.float32, device=device)   # [n][kh][j][c]
    scale = torch.empty(n_total, device=device, dtype=torch.float32)
    bias = torch.empty(n_total, device=device, dtype=torch.float32)
    n0, macs = 0, 0
    for conv, bn in zip(convs, bns):
        cout, cin, kh, kw = conv.weight.shape
        (sh, sw), (ph, pw) = conv.stride, conv.padding
        assert sh == 2 and sw == 2 and cin <= 4 and kh <= STEM_KH and STEM_TOP - ph >= 0 and STEM_LEFT - pw >= 0
        assert kh - ph + STEM_TOP <= STEM_KH and kw - pw + STEM_LEFT <= 8
        w = conv.weight.detach().to(device=device, dtype=torch.float32)
        # tap (ih, iw) reads padded row 2*oh + ih - ph + TOP and padded pixel 2*ow + iw - pw + LEFT
        for ih in range(kh):
            for iw in range(kw):
                wk[n0:n0 + cout, ih - ph + STEM_TOP, iw - pw + STEM_LEFT, :cin] = w[:, :, ih, iw]
        cb = conv.bias.detach().to(device=device, dtype=torch.float32).contiguous() if conv.bias is not None else None
        if bn is not None:
            prm = [t.detach().float().contiguous() for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var)]
            check(lib().amoe_fold_bn(h, ptr(prm[0]), ptr(prm[1]), ptr(prm[2]), ptr(prm[3]), float(bn.eps), ptr(cb), cout,
                                     ptr(scale[n0:]), ptr(bias[n0:]), st), "fold_bn")
        else:
            check(lib().amoe_fold_bn(h, None, None, None, None, 0.0, ptr(cb), cout, ptr(scale[n0:]), ptr(bias[n0:]), st),
                  "fold_bn")
        torch.cuda.current_stream(device).synchronize()
        n0 += cout
        macs += cout * cin * kh * kw
    # [n][k = kh*32 + j*4 + c] -> [k/8][n][8]
    k_total = STEM_KH * 32
    img = wk.reshape(n_total, k_total // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)
    folded = None
    if all(c.weight.shape[1] <= 3 for c in convs):
        # scale folded into the filters; bias = hi + lo (two bf16 values) on the frame's padding channel, which
        # stage_image_stem fills with ones: slots (kh=0, j=0, c=3) and (kh=0, j=1, c=3) are never filter taps
        wf = wk * scale.view(-1, 1, 1, 1)
        assert float(wf[:, 0, 0:2, 3].abs().max()) == 0.0
        hi = bias.to(torch.bfloat16).float()
        wf[:, 0, 0, 3] = hi
        wf[:, 0, 1, 3] = bias - hi
        folded = wf.reshape(n_total, k_total // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)
    return PackedStem(img, scale, bias, couts, n_total, relu, macs, folded)


def stem_forward(ps: PackedStem, x_pad: torch.Tensor, B: int, H: int, W: int,
                 groups: Optional[List[int]] = None) -> List[torch.Tensor]:
    """groups: how many consecutive packed convolutions share one stacked output tensor, e.g. [3, 1] for
    three expert stems ([3B,H/2,W/2,64], the grouped activation layout) + policy conv1 ([B,H/2,W/2,32]).
    Returns one tensor per group."""
    Ho, Wo = H // 2, W // 2
    dev = x_pad.device
    groups = groups or [1] * len(ps.couts)
    assert sum(groups) == len(ps.couts)
    stacked, per_conv, i = [], [], 0
    for n in groups:
        co = ps.couts[i]
        assert all(c == co for c in ps.couts[i:i + n])
        t = torch.empty((n * B, Ho, Wo, co), device=dev, dtype=torch.bfloat16)
        stacked.append(t)
        per_conv += [t[j * B:(j + 1) * B] for j in range(n)]
        i += n
    _stem_launch(ps, x_pad, B, H, W, per_conv)
    return stacked


def stem_pool_supported(H: int, W: int) -> bool:
    import os
    return H % 4 == 0 and W % 4 == 0 and W // 2 <= 128 and os.environ.get("AMOE_STEM_POOL", "1") != "0"


def stem_pool_forward(ps: PackedStem, x_pad: torch.Tensor, B: int, H: int, W: int, n_pool: int, out_pad: int,
                      pooled: Optional[torch.Tensor] = None, rest: Optional[List[torch.Tensor]] = None):
    """First n_pool packed convolutions (64 channels each: expert stems) come back max-pooled as one
    [n_pool*B, H/4+2p, W/4+2p, 64] tensor; the remaining convolutions at full resolution (list).
    `pooled` / `rest` may be preallocated destinations (sub-batch walks reuse them / write into slices)."""
    dev = x_pad.device
    Hp, Wp = H // 4, W // 4
    assert all(c == 64 for c in ps.couts[:n_pool])
    if pooled is None:
        pooled = torch.empty((n_pool * B, Hp + 2 * out_pad, Wp + 2 * out_pad, 64), device=dev, dtype=torch.bfloat16)
    if rest is None:
        rest = [torch.empty((B, H // 2, W // 2, co), device=dev, dtype=torch.bfloat16) for co in ps.couts[n_pool:]]
    n_chunks = ps.n_total // 32
    dst = (C.c_void_p * n_chunks)()
    dst_c = (C.c_int * n_chunks)()
    i = n_pool * 2
    for t, co in zip(rest, ps.couts[n_pool:]):
        for c0 in range(0, co, 32):
            dst[i] = t.data_ptr() + c0 * 2
            dst_c[i] = co
            i += 1
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    fold = ps.w_folded is not None and stem_fold()
    walk_reset(dev, 1)          # the stem writes front to back: the convolution behind it starts at the back
    check(lib().amoe_stem_pool_fwd(ctx(dev), ptr(x_pad), ptr(ps.w_folded if fold else ps.w), None if fold else ptr(ps.scale),
                                   None if fold else ptr(ps.bias), B, H, W, x_pad.shape[2],
                                   STEM_KH, ps.n_total, int(ps.relu), n_pool * 64, ptr(pooled), out_pad, dst, dst_c,
                                   stream_ptr(dev)), "stem_pool_fwd")
    if prof is not None:
        ev1.record()
        prof.append(("stem_pool_kernel", 2.0 * ps.true_macs_per_px * B * (H // 2) * (W // 2), ev0, ev1))
    return pooled, rest


def _stem_launch(ps: PackedStem, x_pad, B, H, W, outs):
    n_chunks = ps.n_total // 32
    dst = (C.c_void_p * n_chunks)()
    dst_c = (C.c_int * n_chunks)()
    i = 0
    for t, co in zip(outs, ps.couts):
        for c0 in range(0, co, 32):
            dst[i] = t.data_ptr() + c0 * 2
            dst_c[i] = co
            i += 1
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(lib().amoe_stem_fwd(ctx(x_pad.device), ptr(x_pad), ptr(ps.w), ptr(ps.scale), ptr(ps.bias), B, H, W,
                              x_pad.shape[2], STEM_KH, ps.n_total, int(ps.relu), dst, dst_c, stream_ptr(x_pad.device)),
          "stem_fwd")
    if prof is not None:
        ev1.record()
        prof.append(("stem_kernel", 2.0 * ps.true_macs_per_px * B * (H // 2) * (W // 2), ev0, ev1))
    return outs


def stem_mode(dtype: torch.dtype, H: Optional[int] = None, W: Optional[int] = None) -> str:
    """How the Cin=3 first-layer convolutions run: 'tc' (bf16 default: one GEMM over raw image rows,
    stem_tc.cu; needs even H, W and W <= 256), 'rowwin' (bf16: expanded row windows through conv_tc.cu - picked
    automatically for frames the 'tc' kernel does not take, e.g. BDD's native 720x1280) or 'simt' (CUDA cores;
    the only choice in fp32 mode).  AMOE_STEM overrides the bf16 choice (debug / A-B switch)."""
    if dtype != torch.bfloat16:
        return "simt"
    m = os.environ.get("AMOE_STEM", "")
    if m in ("tc", "rowwin", "simt"):
        return m
    if H is not None and W is not None and not stem_supported(H, W):
        return "rowwin"
    return "tc"


def l2_chunk_images() -> int:
    """Sub-batch (images per expert) for the L2-resident walk through stem + layer1; 0 disables.
    Off by default: measured on B200 at batch 256 (CUDA-graph replay) chunks of 16/32/64 images gave
    47.5k/48.2k/49.0k frames/s against 50.3k unchunked - the layer1 kernels are bound by shared-memory
    bandwidth, not HBM, so L2 residency buys nothing and the shorter launches pay more tail.
    AMOE_L2_CHUNK=<n> turns it on (kept for larger-L2 / lower-HBM parts and as a tested path)."""
    import os
    return int(os.environ.get("AMOE_L2_CHUNK", "0"))


_WALK = {}   # device index -> direction (0 / 1) of the next tensor-core convolution


def walk_alternate() -> bool:
    """Consecutive tensor-core convolutions walk their tiles in alternating directions (AMOE_WALK_ALT=0: always front to
    back), so each layer starts on the part of its input / residual the previous layer touched last - what is still in L2.
    Results do not depend on the direction (tests run both)."""
    import os
    return os.environ.get("AMOE_WALK_ALT", "1") != "0"


def walk_reset(device, first: int = 0) -> None:
    """Start of a chain (the stem writes front to back): the next convolution walks in direction `first`."""
    _WALK[torch.device(device).index or 0] = first


def _walk_step(device) -> None:
    """Set the direction of the convolution about to be launched on `device`, then flip it for the next one."""
    idx = torch.device(device).index or 0
    d = _WALK.get(idx, 0) if walk_alternate() else 0
    check(lib().amoe_set_walk_reverse(ctx(device), d), "set_walk_reverse")
    _WALK[idx] = d ^ 1


def overlap_tail() -> bool:
    """Run the policy backbone (independent of the experts) on a second stream beside the head / gate kernels
    (AMOE_TAIL_OVERLAP=0 keeps it behind the gate on the main stream)."""
    import os
    return os.environ.get("AMOE_TAIL_OVERLAP", "1") != "0"


def overlap_outputs() -> bool:
    """Fork the full-resolution logit writers onto a side stream (AMOE_OVERLAP=0 keeps one stream)."""
    import os
    return os.environ.get("AMOE_OVERLAP", "1") != "0"


ROWWIN_LEFT = 4     # zero pixels stored left of every image row
ROWWIN_CP = 4       # channels per pixel (3 padded to 4)
ROWWIN_WIN = 16     # pixels per window = 64 bf16 = one 128-byte swizzle row


def rowwin_wpad(W: int, sw: int = 2) -> int:
    Wo = (W - 1) // sw + 1
    need = max(ROWWIN_LEFT + W, (Wo - 1) * sw + ROWWIN_WIN)
    return (need + 7) & ~7


@dataclass
class PackedRowwin:
    """Several small-Cin, stride-2 convolutions that share one input, concatenated on Cout and packed
    for amoe_conv2d_rowwin_fwd (each filter row = one 64-element window of the padded image row)."""
    w: torch.Tensor       # [Cout_total, KH, 64] bf16
    scale: torch.Tensor
    bias: torch.Tensor
    n_conv: int
    cout: int             # per convolution (= split_c)
    kh: int
    kw: int
    sh: int
    sw: int
    ph: int
    pw: int
    relu: bool
    true_k: int


def pack_rowwin(convs, bns, device, relu: bool) -> PackedRowwin:
    c0 = convs[0]
    cout, cin, kh, kw = c0.weight.shape
    sh, sw = c0.stride
    ph, pw = c0.padding
    assert cin <= ROWWIN_CP and kw - pw + ROWWIN_LEFT <= ROWWIN_WIN and ROWWIN_LEFT - pw >= 0
    n = len(convs)
    st, h = stream_ptr(device), ctx(device)
    wbuf = torch.empty((n * cout, kh, 64), device=device, dtype=torch.bfloat16)
    scale = torch.empty(n * cout, device=device, dtype=torch.float32)
    bias = torch.empty(n * cout, device=device, dtype=torch.float32)
    for g, conv in enumerate(convs):
        w = conv.weight.detach().to(device=device, dtype=torch.float32)
        # window position j holds padded column sw*ow + j = original column sw*ow + j - LEFT  ->  kw = j - LEFT + pw
        w2 = w.new_zeros((cout, ROWWIN_WIN, ROWWIN_CP, kh))
        for k in range(kw):
            w2[:, k - pw + ROWWIN_LEFT, :cin, :] = w[:, :, :, k]
        w2 = w2.reshape(cout, 64, kh, 1).contiguous()
        check(lib().amoe_pack_conv_weight(h, ptr(w2), ptr(wbuf[g * cout:]), cout, 64, kh, 1, 64, _cabi.BF16, st),
              "pack_conv_weight")
        cb = conv.bias.detach().to(device=device, dtype=torch.float32).contiguous() if conv.bias is not None else None
        bn = bns[g] if bns is not None else None
        if bn is not None:
            prm = [t.detach().float().contiguous() for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var)]
            check(lib().amoe_fold_bn(h, ptr(prm[0]), ptr(prm[1]), ptr(prm[2]), ptr(prm[3]), float(bn.eps), ptr(cb), cout,
                                     ptr(scale[g * cout:]), ptr(bias[g * cout:]), st), "fold_bn")
        else:
            check(lib().amoe_fold_bn(h, None, None, None, None, 0.0, ptr(cb), cout,
                                     ptr(scale[g * cout:]), ptr(bias[g * cout:]), st), "fold_bn")
        torch.cuda.current_stream(device).synchronize()
    return PackedRowwin(wbuf, scale, bias, n, cout, kh, kw, sh, sw, ph, pw, relu, cin * kh * kw)


def conv2d_rowwin(pc: PackedRowwin, x_pad: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """x_pad: [B,H,Wpad,4] bf16 from image_to_nhwc_padded -> [n_conv*B,Ho,Wo,cout] bf16."""
    Wpad = x_pad.shape[2]
    Ho = (H + 2 * pc.ph - pc.kh) // pc.sh + 1
    Wo = (W + 2 * pc.pw - pc.kw) // pc.sw + 1
    y = torch.empty((pc.n_conv * B, Ho, Wo, pc.cout), device=x_pad.device, dtype=torch.bfloat16)
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(lib().amoe_conv2d_rowwin_fwd(ctx(x_pad.device), ptr(x_pad), ptr(pc.w), ptr(pc.scale), ptr(pc.bias), ptr(y),
                                       B, H, Wpad, ROWWIN_CP, pc.n_conv * pc.cout, pc.cout, pc.kh, pc.sh, pc.sw, pc.ph,
                                       Ho, Wo, int(pc.relu), stream_ptr(x_pad.device)), "conv2d_rowwin_fwd")
    if prof is not None:
        ev1.record()
        prof.append(("conv_tc_kernel", 2.0 * pc.true_k * pc.cout * pc.n_conv * B * Ho * Wo, ev0, ev1))
    return y


@dataclass
class PackedConv:
    """Packed weights of G same-shape convolutions (+ folded BN) living on one device."""
    w: torch.Tensor        # [G*Cout, KH, KW, Cin_k] (dtype)
    scale: torch.Tensor    # [G*Cout] fp32
    bias: torch.Tensor     # [G*Cout] fp32
    G: int
    cin: int               # channels the kernel sees (after padding / pairing)
    cout: int
    kh: int
    kw: int
    sh: int
    sw: int
    ph: int
    pw: int
    relu: bool
    pair_w: bool = False   # input is read through the [N,H,W/2,2C] pixel-pair view
    meta: dict = field(default_factory=dict)


def pack_conv(convs, bns, dtype: torch.dtype, device, relu: bool, cin_pad: Optional[int] = None,
              allow_pair: bool = True) -> PackedConv:
    """Pack G nn.Conv2d (+ optional eval-mode nn.BatchNorm2d each) for amoe_conv2d_fwd.

    BN (running stats) and the conv bias are folded into per-channel scale/bias applied
    in the conv epilogue in fp32 (torchvision BasicBlock bn1/bn2; trajectory_head.py:8-24).
    """
    G = len(convs)
    c0 = convs[0]
    cout, cin, kh, kw = c0.weight.shape
    sh, sw = c0.stride
    ph, pw = c0.padding
    st = stream_ptr(device)
    h = ctx(device)
    # 3x3/s2 convolutions with 32 input channels are re-expressed over pixel pairs so the
    # implicit GEMM sees 64 contiguous channels per tap ("pair view").
    pair = (allow_pair and dtype == torch.bfloat16 and cin == 32 and sw == 2 and kw == 3 and pw == 1)
    if pair:
        cin_k, kw_k, sw_k, pw_k = 64, 2, 1, 1
    else:
        cin_k, kw_k, sw_k, pw_k = (cin_pad or cin), kw, sw, pw
    wbuf = torch.empty((G * cout, kh, kw_k, cin_k), device=device, dtype=dtype)
    scale = torch.empty(G * cout, device=device, dtype=torch.float32)
    bias = torch.empty(G * cout, device=device, dtype=torch.float32)
    for g, conv in enumerate(convs):
        w = conv.weight.detach().to(device=device, dtype=torch.float32).contiguous()
        if pair:
            # w'[o, par*32+c, kh, j]: (j=0,par=1)=kw0, (j=1,par=0)=kw1, (j=1,par=1)=kw2, (j=0,par=0)=0
            w2 = w.new_zeros((cout, 64, kh, 2))
            w2[:, 32:, :, 0] = w[:, :, :, 0]
            w2[:, :32, :, 1] = w[:, :, :, 1]
            w2[:, 32:, :, 1] = w[:, :, :, 2]
            w = w2.contiguous()
        check(lib().amoe_pack_conv_weight(h, ptr(w), ptr(wbuf[g * cout:]), cout, w.shape[1], kh, kw_k, cin_k,
                                          dtype_code(dtype), st), "pack_conv_weight")
        cb = conv.bias.detach().to(device=device, dtype=torch.float32).contiguous() if conv.bias is not None else None
        bn = bns[g] if bns is not None else None
        if bn is not None:
            gmm = bn.weight.detach().to(device=device, dtype=torch.float32).contiguous()
            bta = bn.bias.detach().to(device=device, dtype=torch.float32).contiguous()
            mean = bn.running_mean.detach().to(device=device, dtype=torch.float32).contiguous()
            var = bn.running_var.detach().to(device=device, dtype=torch.float32).contiguous()
            check(lib().amoe_fold_bn(h, ptr(gmm), ptr(bta), ptr(mean), ptr(var), float(bn.eps), ptr(cb), cout,
                                     ptr(scale[g * cout:]), ptr(bias[g * cout:]), st), "fold_bn")
        else:
            check(lib().amoe_fold_bn(h, None, None, None, None, 0.0, ptr(cb), cout,
                                     ptr(scale[g * cout:]), ptr(bias[g * cout:]), st), "fold_bn")
        # the temporaries above must outlive the async kernels that read them
        torch.cuda.current_stream(device).synchronize()
    meta = {"true_k": cin * kh * kw}
    if dtype == torch.float32 and f32_tc() and sh == sw and ph == pw and cin_k == cin and \
            bool(lib().amoe_conv2d_f32tc_supported(64, 64, cin, cout, kh, kw, 1)):
        # fp32 mode on the tensor cores: six-term split of the weights (csrc/conv_tc.cu, "split operands")
        wsplit = torch.empty((G * cout, kh, kw, 6, cin), device=device, dtype=torch.bfloat16)
        for g, conv in enumerate(convs):
            w = conv.weight.detach().to(device=device, dtype=torch.float32).contiguous()
            check(lib().amoe_pack_conv_weight_split6(h, ptr(w), ptr(wsplit[g * cout:]), cout, cin, kh, kw, 0, st), "pack_conv_weight_split6")
            torch.cuda.current_stream(device).synchronize()
        meta["w_split"] = wsplit
    return PackedConv(wbuf, scale, bias, G, cin_k, cout, kh, kw_k, sh, sw_k, ph, pw_k, relu, pair, meta=meta)


def f32_tc() -> bool:
    """fp32 ("parity") mode: convolutions with Cin % 64 == 0 run fp32-accurate on the bf16 tensor cores (three-way split
    operands, csrc/conv_tc.cu); AMOE_F32_TC=0 keeps them on the CUDA-core kernel."""
    return os.environ.get("AMOE_F32_TC", "1") != "0"


def split3(x: torch.Tensor) -> torch.Tensor:
    """[..., C] fp32 -> [..., 3C] bf16 = (x1 | x2 | x3): the three-way split of the fp32-accurate tensor-core kernels."""
    Cc = x.shape[-1]
    out = torch.empty(x.shape[:-1] + (3 * Cc,), device=x.device, dtype=torch.bfloat16)
    check(lib().amoe_split3_bf16(ctx(x.device), ptr(x), ptr(out), x.numel() // Cc, Cc, stream_ptr(x.device)), "split3_bf16")
    return out


def use_flat() -> bool:
    """bf16 path: 3x3/s1 convolutions of the 64/128-channel stages through the halo-reuse kernel on
    physically padded activations (csrc/conv_flat.cu).  AMOE_FLAT=0 keeps the per-tap TMA kernel."""
    import os
    return os.environ.get("AMOE_FLAT", "1") != "0"


def conv2d(pc: PackedConv, x: torch.Tensor, B: int, H: int, W: int, residual: Optional[torch.Tensor] = None,
           x_shared: bool = False, impl: int = 0, relu: Optional[bool] = None, in_pad: int = 0, out_pad: int = 0,
           zero_border: bool = False) -> torch.Tensor:
    """x: [G*B,H(+2*in_pad),W(+2*in_pad),Cin] NHWC (or [B,...] if x_shared) -> [G*B,Ho(+2*out_pad),Wo(+2*out_pad),Cout].
    H, W name the interior size.  zero_border: the padded output is allocated zero-filled."""
    dtype = x.dtype
    if pc.pair_w:
        # output geometry of the original 3x3/s2/p1 conv; the kernel sees width W/2, KW=2, stride_w=1
        Ho, Wo = (H + 2 * pc.ph - pc.kh) // pc.sh + 1, W // 2
        Wk = W // 2
    else:
        Ho = (H + 2 * pc.ph - pc.kh) // pc.sh + 1
        Wo = (W + 2 * pc.pw - pc.kw) // pc.sw + 1
        Wk = W
    shape = (pc.G * B, Ho + 2 * out_pad, Wo + 2 * out_pad, pc.cout)
    # out_pad == 1 on the tcgen05 path: the kernel writes the zero border itself (a memset of the whole tensor cost 60 us per
    # 175 MB layer2 activation); other cases keep the zero-filled allocation
    tc_border = out_pad == 1 and dtype == torch.bfloat16 and impl != 1
    y = (torch.zeros if (zero_border and out_pad and not tc_border) else torch.empty)(shape, device=x.device, dtype=dtype)
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _walk_step(x.device)
    wsplit = pc.meta.get("w_split")
    f32tc = (dtype == torch.float32 and wsplit is not None and impl == 0 and in_pad == 0 and out_pad == 0 and f32_tc() and
             bool(lib().amoe_conv2d_f32tc_supported(H, Wk, pc.cin, pc.cout, pc.kh, pc.kw, pc.sh)))
    if f32tc:
        xs = split3(x.contiguous())       # named: must outlive the launch
        check(lib().amoe_conv2d_fwd_f32tc_grouped(ctx(x.device), ptr(xs), ptr(wsplit), ptr(pc.scale), ptr(pc.bias), ptr(residual),
                                                  ptr(y), pc.G, int(x_shared), B, H, Wk, pc.cin, pc.cout, pc.kh, pc.kw, pc.sh,
                                                  pc.ph, Ho, Wo, int(pc.relu if relu is None else relu), stream_ptr(x.device)),
              "conv2d_fwd_f32tc_grouped")
        del xs
    else:
        check(lib().amoe_conv2d_fwd(ctx(x.device), ptr(x), ptr(pc.w), ptr(pc.scale), ptr(pc.bias), ptr(residual), ptr(y),
                                    pc.G, int(x_shared), B, H, Wk, pc.cin, pc.cout, pc.kh, pc.kw, pc.sh, pc.sw,
                                    pc.ph, pc.pw, Ho, Wo, int(pc.relu if relu is None else relu), dtype_code(dtype),
                                    impl, in_pad, out_pad, stream_ptr(x.device)), "conv2d_fwd")
    if prof is not None:
        ev1.record()
        tc = f32tc or (dtype == torch.bfloat16 and impl != 1 and lib().amoe_conv2d_tc_supported(H + 2 * in_pad, Wk + 2 * in_pad, pc.cin, pc.cout, pc.sh, pc.sw))
        macs = pc.meta.get("true_k", pc.kh * pc.kw * pc.cin) * pc.cout * pc.G * B * Ho * Wo
        prof.append(("conv_tc_kernel" if tc else "conv2d_simt_kernel", 2.0 * macs, ev0, ev1))
    return y


def dual_entry() -> bool:
    """Stage-entry conv1 (3x3/s2) and the block's 1x1/s2 downsample as one dual launch (AMOE_DUAL=0: two launches)."""
    return os.environ.get("AMOE_DUAL", "1") != "0"


def conv2d_dual(pc: PackedConv, pd: PackedConv, x: torch.Tensor, B: int, H: int, W: int, in_pad: int = 0, out_pad: int = 0):
    """conv (pc: KxK/stride s/pad p, + BN + ReLU) and the 1x1/stride s/pad 0 convolution pd (+ BN) over the same input in
    one launch -> (y, y2), both [G*B,Ho(+2*out_pad),Wo(+2*out_pad),Cout] with a zero border when out_pad."""
    assert x.dtype == torch.bfloat16 and pc.sh == pc.sw == pd.sh == pd.sw and pc.ph == pc.pw and not pc.pair_w
    assert pd.kh == pd.kw == 1 and pd.ph == pd.pw == 0 and pd.cout == pc.cout and pd.cin == pc.cin and pd.G == pc.G
    Ho = (H + 2 * pc.ph - pc.kh) // pc.sh + 1
    Wo = (W + 2 * pc.pw - pc.kw) // pc.sw + 1
    assert Ho == (H - 1) // pd.sh + 1 and Wo == (W - 1) // pd.sw + 1
    shape = (pc.G * B, Ho + 2 * out_pad, Wo + 2 * out_pad, pc.cout)
    assert out_pad in (0, 1)
    # y: zero border written by the kernel (out_pad == 1); y2 is only ever read as a residual, whose border is ignored
    y, y2 = torch.empty(shape, device=x.device, dtype=x.dtype), torch.empty(shape, device=x.device, dtype=x.dtype)
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _walk_step(x.device)
    check(lib().amoe_conv2d_dual_fwd(ctx(x.device), ptr(x), ptr(pc.w), ptr(pc.scale), ptr(pc.bias), ptr(y), ptr(pd.w), ptr(pd.scale),
                                     ptr(pd.bias), ptr(y2), pc.G, B, H, W, pc.cin, pc.cout, pc.kh, pc.kw, pc.sh, pc.ph, Ho, Wo,
                                     int(pc.relu), int(pd.relu), in_pad, out_pad, stream_ptr(x.device)), "conv2d_dual_fwd")
    if prof is not None:
        ev1.record()
        macs = (pc.kh * pc.kw + 1) * pc.cin * pc.cout * pc.G * B * Ho * Wo
        prof.append(("conv_tc_kernel", 2.0 * macs, ev0, ev1))
    return y, y2


def dual_supported(pc: PackedConv, pd: Optional[PackedConv], H: int, W: int, in_pad: int, dtype: torch.dtype) -> bool:
    return (pd is not None and dtype == torch.bfloat16 and dual_entry() and not pc.pair_w and pc.sh == pc.sw and pc.ph == pc.pw
            and pd.kh == 1 and pd.kw == 1 and pd.sh == pc.sh and pd.sw == pc.sw and pd.ph == 0 and pd.pw == 0
            and bool(lib().amoe_conv2d_tc_supported(H + 2 * in_pad, W + 2 * in_pad, pc.cin, pc.cout, pc.sh, pc.sw)))


def flat_supported(pc: PackedConv, H: int, W: int, dtype: torch.dtype) -> bool:
    return (dtype == torch.bfloat16 and pc.kh == 3 and pc.kw == 3 and pc.sh == 1 and pc.sw == 1 and pc.ph == 1
            and pc.pw == 1 and not pc.pair_w and bool(lib().amoe_conv3x3_flat_supported(H, W, pc.cin, pc.cout)))


def conv3x3_flat(pc: PackedConv, x_pad: torch.Tensor, B: int, H: int, W: int,
                 residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 out_group_images: int = 0) -> torch.Tensor:
    """x_pad: [G*B,H+2,W+2,Cin] bf16 with zero border -> [G*B,H+2,W+2,Cout] with zero border.
    `out` (+ out_group_images): write the B images of every expert group into a larger grouped tensor
    whose groups are out_group_images apart; `out` starts at this sub-batch's first image of group 0."""
    y = out if out is not None else torch.empty((pc.G * B, H + 2, W + 2, pc.cout), device=x_pad.device, dtype=x_pad.dtype)
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _walk_step(x_pad.device)
    check(lib().amoe_conv3x3_flat_fwd_strided(ctx(x_pad.device), ptr(x_pad), ptr(pc.w), ptr(pc.scale), ptr(pc.bias),
                                              ptr(residual), ptr(y), pc.G, B, H, W, pc.cin, pc.cout, int(pc.relu),
                                              int(out_group_images), 0, stream_ptr(x_pad.device)),
          "conv3x3_flat_fwd")
    if prof is not None:
        ev1.record()
        prof.append((f"conv3x3_flat_kernel<{pc.cout}>", 2.0 * 9 * pc.cin * pc.cout * pc.G * B * H * W, ev0, ev1))
    return y


def maxpool3x3s2(x: torch.Tensor, out_pad: int = 0) -> torch.Tensor:
    NB, H, W, Cc = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((NB, Ho + 2 * out_pad, Wo + 2 * out_pad, Cc), device=x.device, dtype=x.dtype)
    check(lib().amoe_maxpool3x3s2_fwd(ctx(x.device), ptr(x), ptr(y), NB, H, W, Cc, dtype_code(x.dtype), out_pad,
                                      stream_ptr(x.device)), "maxpool3x3s2_fwd")
    return y


def head1x1_pool(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, pooled: torch.Tensor, pooled_off: int):
    """x: [B,h,w,Cin]; w: [N,Cin] fp32; writes pooled[:, off:off+N]; returns low [B,h,w,N] fp32."""
    B, h, wd, Cin = x.shape
    N = w.shape[0]
    low = torch.empty((B, h, wd, N), device=x.device, dtype=torch.float32)
    pview = pooled[:, pooled_off:]
    check(lib().amoe_head1x1_pool_fwd(ctx(x.device), ptr(x), ptr(w), ptr(b), ptr(low), ptr(pview), pooled.shape[1],
                                      B, h * wd, Cin, N, dtype_code(x.dtype), stream_ptr(x.device)), "head1x1_pool_fwd")
    return low


def upsample_bilinear_nchw(low: torch.Tensor, H: int, W: int, dtype: torch.dtype) -> torch.Tensor:
    B, h, w, Cc = low.shape
    out = torch.empty((B, Cc, H, W), device=low.device, dtype=dtype)
    check(lib().amoe_upsample_bilinear_nchw_fwd(ctx(low.device), ptr(low), ptr(out), B, h, w, Cc, H, W,
                                                dtype_code(dtype), stream_ptr(low.device)), "upsample_bilinear_nchw_fwd")
    return out


def mean_hw_nchw(x: torch.Tensor) -> torch.Tensor:
    B, Cc, H, W = x.shape
    x = x.contiguous()
    out = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    check(lib().amoe_mean_hw_nchw_fwd(ctx(x.device), ptr(x), ptr(out), B, Cc, H * W, dtype_code(x.dtype),
                                      stream_ptr(x.device)), "mean_hw_nchw_fwd")
    return out


def gate(state, pooled, params, n_ch, ctx_dim, hidden, temperature, mode=0, params_bf16=None, ext_features=None):
    """Fused context extractor + expert extractors + gating network (gate.cu).
    params_bf16: bf16 copy of `params` -> tensor-core (TF32) variant for bf16 inference at B >= 16.
    ext_features: [E,B,256] fp32 - features computed outside the kernel for the experts whose n_ch entry is 0."""
    dev = state.device
    B, E = state.shape[0], len(n_ch)
    f32 = dict(device=dev, dtype=torch.float32)
    context = torch.empty((B, ctx_dim), **f32)
    weights = torch.empty((B, E), **f32)
    logits = torch.empty((B, E), **f32)
    if mode & 1:
        features = processed = combined = None
    else:
        features = torch.empty((E, B, 256), **f32)
        processed = torch.empty((E, B, 256), **f32)
        combined = torch.empty((B, 256), **f32)
    arr = (C.c_int * E)(*n_ch)
    if params_bf16 is not None:
        assert params_bf16.dtype == torch.bfloat16 and params_bf16.numel() == params.numel()
    if ext_features is not None:
        assert ext_features.dtype == torch.float32 and tuple(ext_features.shape) == (E, B, 256) and ext_features.is_contiguous()
    check(lib().amoe_gate_fwd_ex2(ctx(dev), ptr(state), ptr(pooled), ptr(params), ptr(params_bf16), params.numel(), B, E,
                                  arr, ctx_dim, hidden, float(temperature), mode, ptr(ext_features), ptr(context), ptr(features),
                                  ptr(processed), ptr(logits), ptr(weights), ptr(combined), stream_ptr(dev)), "gate_fwd")
    return dict(context=context, features=features, processed=processed, gate_logits=logits, weights=weights,
                combined=combined)


def mlp_tc(dtype: torch.dtype) -> bool:
    """bf16 inference mode runs the gate / policy-head MLPs 16 frames per CTA on mma.sync TF32 with bf16
    weights (AMOE_MLP_TC=0 keeps the fp32 CUDA-core kernels; fp32 mode always uses those)."""
    return dtype == torch.bfloat16 and os.environ.get("AMOE_MLP_TC", "1") != "0"


def mean_hw_nhwc(x):
    """[B,h,w,C] bf16 -> [B,C] fp32 mean over the pixels (deterministic order)."""
    B, h, w, Cc = x.shape
    out = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    check(lib().amoe_mean_hw_nhwc_fwd(ctx(x.device), ptr(x), ptr(out), B, h * w, Cc, dtype_code(x.dtype),
                                      stream_ptr(x.device)), "mean_hw_nhwc_fwd")
    return out


def policy_head(x, cvec, params, backbone_dim, ctx_dim, hidden, horizon, params_bf16=None):
    """x: [B,h,w,Cf] conv4 output; cvec: [B,ctx_dim] fp32 or None.
    params_bf16: bf16 copy of `params` -> tensor-core (TF32) variant for bf16 inference at B >= 16; the
    pooling then runs as its own grid-wide kernel (one CTA per frame) instead of inside the head's 16 CTAs."""
    B, h, w, Cf = x.shape
    dev = x.device
    wp = torch.empty((B, 2 * horizon), device=dev, dtype=torch.float32)
    spd = torch.empty((B, horizon), device=dev, dtype=torch.float32)
    if params_bf16 is not None and B >= 16:
        assert params_bf16.dtype == torch.bfloat16 and params_bf16.numel() == params.numel()
        if x.dtype == torch.bfloat16 and Cf % 8 == 0:     # (already pooled by TrajectoryPolicy.backbone_features otherwise)
            x = mean_hw_nhwc(x).view(B, 1, 1, Cf)
            h = w = 1
    else:
        params_bf16 = None
    check(lib().amoe_policy_head_fwd_ex(ctx(dev), ptr(x), ptr(cvec), ptr(params), params.numel(), B, h * w, Cf,
                                        backbone_dim, ctx_dim, hidden, horizon, dtype_code(x.dtype), ptr(params_bf16),
                                        ptr(wp), ptr(spd), stream_ptr(dev)), "policy_head_fwd")
    return wp, spd


def hungarian_cost(logits, boxes, tgt_boxes, tgt_labels, n_tgt, w_class, w_bbox, w_giou):
    B, Q, Cc = logits.shape
    D = boxes.shape[2]
    Nmax = tgt_boxes.shape[1]
    cost = torch.empty((B, Q, Nmax), device=logits.device, dtype=torch.float32)
    check(lib().amoe_hungarian_cost_fwd(ctx(logits.device), ptr(logits), ptr(boxes), ptr(tgt_boxes), ptr(tgt_labels),
                                        ptr(n_tgt), ptr(cost), B, Q, Cc, D, Nmax, float(w_class), float(w_bbox),
                                        float(w_giou), stream_ptr(logits.device)), "hungarian_cost_fwd")
    return cost


def lsap_batched(cost_host: torch.Tensor, n_tgt_host: torch.Tensor, n_threads: int = 8):
    """cost_host: [B,Q,Nmax] fp32 CPU tensor; returns (rows, cols, n_match) CPU tensors."""
    B, Q, Nmax = cost_host.shape
    K = min(Q, Nmax)
    rows = torch.zeros((B, K), dtype=torch.int64)
    cols = torch.zeros((B, K), dtype=torch.int64)
    nm = torch.zeros((B,), dtype=torch.int32)
    cost_host = cost_host.contiguous()
    n_tgt_host = n_tgt_host.to(torch.int32).contiguous()
    rc = lib().amoe_lsap_batched_host(ptr(cost_host), ptr(n_tgt_host), B, Q, Nmax, ptr(rows), ptr(cols), ptr(nm),
                                      n_threads)
    if rc == -2 or rc == -3:
        # scipy.optimize.linear_sum_assignment raises ValueError for NaN/-inf or infeasible matrices
        raise ValueError(lib().amoe_last_error().decode())
    check(rc, "lsap_batched_host")
    return rows, cols, nm


# ---- input staging from camera bytes (csrc/stage_u8.cu; reference inference/run_automoe.py:25-31) ----
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _f3(v):
    a = torch.as_tensor(v, dtype=torch.float32).reshape(3)      # the fp32 rounding T.Normalize applies to mean/std
    return (C.c_float * 3)(*[float(x) for x in a])


def _check_u8_hwc(img: torch.Tensor):
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] != 3:
        raise ValueError(f"expected uint8 frames [B,H,W,3] (HWC RGB), got {img.dtype} {tuple(img.shape)}")
    if not img.is_cuda:
        raise RuntimeError("automoe_b200 has no CPU path: move the frames to a CUDA (sm_100a) device")


def stage_u8_stem(img_u8: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """uint8 [B,H,W,3] -> the frame amoe_stem_fwd reads ([B,H+6,Wpad,4] bf16, normalised, 4th channel 1.0):
    ToTensor + Normalize + stage_image_stem as one kernel."""
    _check_u8_hwc(img_u8)
    img_u8 = img_u8.contiguous()
    B, H, W, _ = img_u8.shape
    out = torch.empty((B, H + 6, stem_wpad(W), 4), device=img_u8.device, dtype=torch.bfloat16)
    check(lib().amoe_stage_u8_hwc_fwd(ctx(img_u8.device), ptr(img_u8), ptr(out), B, H, W, STEM_LEFT, out.shape[2], STEM_TOP,
                                      H + 6, _f3(mean), _f3(std), 1.0, stream_ptr(img_u8.device)), "stage_u8_hwc_fwd")
    return out


def normalize_u8_nchw(img_u8: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """uint8 [B,H,W,3] -> fp32 [B,3,H,W], bit-identical to T.ToTensor() + T.Normalize(mean, std)."""
    _check_u8_hwc(img_u8)
    img_u8 = img_u8.contiguous()
    B, H, W, _ = img_u8.shape
    out = torch.empty((B, 3, H, W), device=img_u8.device, dtype=torch.float32)
    check(lib().amoe_normalize_u8_hwc_to_nchw_fwd(ctx(img_u8.device), ptr(img_u8), ptr(out), B, H, W, _f3(mean), _f3(std),
                                                  stream_ptr(img_u8.device)), "normalize_u8_hwc_to_nchw_fwd")
    return out


_PIL_PRECISION_BITS = 32 - 8 - 2
_resize_tables: dict = {}


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Fixed-point taps of Pillow's 8-bit bilinear resize along one axis (Pillow src/libImaging/Resample.c:
    precompute_coeffs with the bilinear filter (support 1, widened by the scale when shrinking = antialias) and
    normalize_coeffs_8bpc).  Returns (bounds int32 [out,2], coeffs int32 [out,ksize], ksize).  Python floats are
    the C doubles of the original; int() truncates like the C cast."""
    import math
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = torch.zeros((out_size, 2), dtype=torch.int32)
    coeffs = torch.zeros((out_size, ksize), dtype=torch.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            k.append(1.0 - a if a < 1.0 else 0.0)
        ww = 0.0                        # accumulated left to right in double, like the C loop
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            coeffs[xx, x] = int(-0.5 + v * (1 << _PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PIL_PRECISION_BITS))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, coeffs, ksize


def resize_u8_bilinear(img_u8: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """uint8 [B,H,W,3] -> uint8 [B,out_h,out_w,3]: Pillow's Image.resize(BILINEAR) (horizontal pass, then vertical pass,
    8-bit rounding after each), the T.Resize of inference/run_automoe.py:27.  Same size returns the input (Pillow copies)."""
    _check_u8_hwc(img_u8)
    B, H, W, _ = img_u8.shape
    if (H, W) == (out_h, out_w):
        return img_u8
    dev = img_u8.device
    x = img_u8.contiguous()

    def table(n_in, n_out):
        key = (n_in, n_out, dev.index)
        t = _resize_tables.get(key)
        if t is None:
            b, c, ks = pil_bilinear_coeffs(n_in, n_out)
            t = _resize_tables[key] = (b.to(dev), c.to(dev), ks)
        return t

    if W != out_w:
        b, c, ks = table(W, out_w)
        y = torch.empty((B, H, out_w, 3), device=dev, dtype=torch.uint8)
        check(lib().amoe_resample_u8_fwd(ctx(dev), ptr(x), ptr(y), ptr(b), ptr(c), ks, B * H, W, out_w, 3, stream_ptr(dev)),
              "resample_u8_fwd(horizontal)")
        x = y
    if H != out_h:
        b, c, ks = table(H, out_h)
        y = torch.empty((B, out_h, out_w, 3), device=dev, dtype=torch.uint8)
        check(lib().amoe_resample_u8_fwd(ctx(dev), ptr(x), ptr(y), ptr(b), ptr(c), ks, B, H, out_h, out_w * 3, stream_ptr(dev)),
              "resample_u8_fwd(vertical)")
        x = y
    return x
