"""Differentiable building blocks of the gating / policy training step.

Each torch.autograd.Function below is a thin shim over the C-ABI training kernels
(include/automoe_b200.h, csrc/train_mlp.cu, csrc/train_conv.cu): forward and backward both run
hand-written sm_100a kernels in fp32; torch only owns the buffers and strings the graph together,
so `loss.backward()`, `clip_grad_norm_`, `AdamW` and `DistributedDataParallel` of the reference's
trainer (training/train_gating_network.py:76-117) work unchanged on the drop-in modules.

There is no torch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _cabi
from .._cabi import check, ctx, lib
from .._ops import ptr, stream_ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("automoe_b200 has no CPU path: training tensors must live on a CUDA (sm_100a) device")
    return t.detach().to(torch.float32).contiguous()


def next_seed() -> int:
    """Dropout seed drawn from torch's CPU generator: reproducible under torch.manual_seed, no device sync."""
    return int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFFFFFFFFFF


# Per-step part of the dropout key on the device (a one-element int64 tensor) while a training step is captured as a CUDA
# graph: the host seed drawn per call is baked into the graph, the device part advances on every replay (amoe_train_tick),
# so every replay draws new masks.  None: eager mode, the key is the host seed alone.
_DEVICE_SEED: Optional[torch.Tensor] = None


class device_dropout_seed:
    """Context manager: dropout layers executed inside key their masks by (host seed + *seed_dev)."""

    def __init__(self, seed_dev: Optional[torch.Tensor]):
        self.seed_dev = seed_dev

    def __enter__(self):
        global _DEVICE_SEED
        self.prev, _DEVICE_SEED = _DEVICE_SEED, self.seed_dev
        return self

    def __exit__(self, *exc):
        global _DEVICE_SEED
        _DEVICE_SEED = self.prev
        return False


# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    """y = Dropout_p(ReLU?(x W^T + b)) — nn.Linear [+ nn.ReLU [+ nn.Dropout]] fused."""

    @staticmethod
    def forward(ctx_, x, W, b, relu: bool, drop_p: float, seed: int):
        x2, W2 = _f32c(x), _f32c(W)
        b2 = _f32c(b) if b is not None else None
        B, in_dim = x2.shape
        out_dim = W2.shape[0]
        y = torch.empty((B, out_dim), device=x2.device, dtype=torch.float32)
        if drop_p > 0.0 and _DEVICE_SEED is not None:
            check(lib().amoe_linear_fwd_dseed(ctx(x2.device), ptr(x2), in_dim, ptr(W2), ptr(b2), ptr(y), out_dim, B, in_dim, out_dim,
                                              int(relu), float(drop_p), seed, ptr(_DEVICE_SEED), stream_ptr(x2.device)),
                  "linear_fwd_dseed")
        else:
            check(lib().amoe_linear_fwd(ctx(x2.device), ptr(x2), in_dim, ptr(W2), ptr(b2), ptr(y), out_dim, B, in_dim, out_dim,
                                        int(relu), float(drop_p), seed, stream_ptr(x2.device)), "linear_fwd")
        ctx_.save_for_backward(x2, W2, y if relu else None)
        ctx_.cfg = (relu, float(drop_p), b is not None)
        return y

    @staticmethod
    def backward(ctx_, dy):
        x2, W2, y = ctx_.saved_tensors
        relu, drop_p, has_b = ctx_.cfg
        dy = _f32c(dy)
        B, in_dim = x2.shape
        out_dim = W2.shape[0]
        need_x, need_W, need_b = ctx_.needs_input_grad[0], ctx_.needs_input_grad[1], has_b and ctx_.needs_input_grad[2]
        dev = x2.device
        dx = torch.empty_like(x2) if need_x else None
        dW = torch.empty_like(W2) if need_W else None
        db = torch.empty(out_dim, device=dev, dtype=torch.float32) if need_b else None
        g_tmp = torch.empty((B, out_dim), device=dev, dtype=torch.float32) if relu else None
        check(lib().amoe_linear_bwd(ctx(dev), ptr(dy), out_dim, ptr(y), out_dim, ptr(x2), in_dim, ptr(W2), ptr(g_tmp),
                                    ptr(dx), in_dim, ptr(dW), ptr(db), B, in_dim, out_dim, int(relu), drop_p,
                                    stream_ptr(dev)), "linear_bwd")
        return dx, dW, db, None, None, None


def linear(x, lin: nn.Linear, relu: bool = False, drop_p: float = 0.0) -> torch.Tensor:
    seed = next_seed() if drop_p > 0.0 else 0
    return _Linear.apply(x, lin.weight, lin.bias, relu, drop_p, seed)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx_, x, gamma, beta, eps: float):
        x2, g2, b2 = _f32c(x), _f32c(gamma), _f32c(beta)
        B, D = x2.shape
        dev = x2.device
        y = torch.empty_like(x2)
        mean = torch.empty(B, device=dev, dtype=torch.float32)
        rstd = torch.empty(B, device=dev, dtype=torch.float32)
        check(lib().amoe_layernorm_fwd(ctx(dev), ptr(x2), ptr(g2), ptr(b2), ptr(y), ptr(mean), ptr(rstd), B, D, float(eps),
                                       stream_ptr(dev)), "layernorm_fwd")
        ctx_.save_for_backward(x2, g2, mean, rstd)
        return y

    @staticmethod
    def backward(ctx_, dy):
        x2, g2, mean, rstd = ctx_.saved_tensors
        dy = _f32c(dy)
        B, D = x2.shape
        dev = x2.device
        dx = torch.empty_like(x2) if ctx_.needs_input_grad[0] else None
        need_p = ctx_.needs_input_grad[1] or ctx_.needs_input_grad[2]
        dg = torch.empty(D, device=dev, dtype=torch.float32) if need_p else None
        db = torch.empty(D, device=dev, dtype=torch.float32) if need_p else None
        check(lib().amoe_layernorm_bwd(ctx(dev), ptr(dy), ptr(x2), ptr(g2), ptr(mean), ptr(rstd), ptr(dx), ptr(dg), ptr(db),
                                       B, D, stream_ptr(dev)), "layernorm_bwd")
        return dx, dg, db, None


def layer_norm(x, ln: nn.LayerNorm) -> torch.Tensor:
    return _LayerNorm.apply(x, ln.weight, ln.bias, ln.eps)


class _GateCombine(torch.autograd.Function):
    """(logits [B,E], processed [E,B,P]) -> (weights [B,E], combined [B,P])  — gating_network.py:157-165."""

    @staticmethod
    def forward(ctx_, logits, processed, temperature: float, use_softmax: bool = True):
        lg, pr = _f32c(logits), _f32c(processed)
        E, B, P = pr.shape
        dev = lg.device
        weights = torch.empty((B, E), device=dev, dtype=torch.float32)
        combined = torch.empty((B, P), device=dev, dtype=torch.float32)
        check(lib().amoe_gate_combine_fwd_ex(ctx(dev), ptr(lg), ptr(pr), B * P, P, float(temperature), int(bool(use_softmax)),
                                             ptr(weights), ptr(combined), B, E, P, stream_ptr(dev)), "gate_combine_fwd")
        ctx_.save_for_backward(weights, pr, lg)
        ctx_.T = float(temperature)
        ctx_.use_softmax = bool(use_softmax)
        return weights, combined

    @staticmethod
    def backward(ctx_, dweights, dcombined):
        weights, pr, lg = ctx_.saved_tensors
        E, B, P = pr.shape
        dev = pr.device
        dw = _f32c(dweights) if dweights is not None else None
        dc = _f32c(dcombined) if dcombined is not None else None
        dlogits = torch.empty((B, E), device=dev, dtype=torch.float32)
        dproc = torch.empty_like(pr)
        check(lib().amoe_gate_combine_bwd_ex(ctx(dev), ptr(dc), ptr(dw), ptr(weights), ptr(pr), B * P, P, ctx_.T,
                                             None if ctx_.use_softmax else ptr(lg), ptr(dlogits), ptr(dproc), B * P, B, E, P,
                                             stream_ptr(dev)), "gate_combine_bwd")
        return dlogits, dproc, None, None


def gate_combine(logits, processed: List[torch.Tensor], temperature: float,
                 use_softmax: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    return _GateCombine.apply(logits, torch.stack(processed, dim=0), temperature, use_softmax)


# ------------------------------------------------------------------------------------------------
def _pack_w(weight: torch.Tensor, cin_pad: int) -> torch.Tensor:
    """nn.Conv2d weight [Cout,Cin,KH,KW] -> packed [Cout,KH,KW,cin_pad] fp32 (kernel layout)."""
    Cout, Cin, KH, KW = weight.shape
    w = _f32c(weight)
    out = torch.empty((Cout, KH, KW, cin_pad), device=w.device, dtype=torch.float32)
    check(lib().amoe_pack_conv_weight(ctx(w.device), ptr(w), ptr(out), Cout, Cin, KH, KW, cin_pad, _cabi.F32,
                                      stream_ptr(w.device)), "pack_conv_weight")
    return out


BN_NONE, BN_BATCH, BN_RUNNING = 0, 1, 2


# ---- fp32-accurate convolutions on the bf16 tensor cores (csrc/conv_tc.cu: split operands) ----
def train_tc() -> bool:
    """Training convolutions with Cin % 64 == 0 run on tcgen05 with split bf16 operands (fp32-accurate: the 1e-4 parity
    gates hold); AMOE_TRAIN_TC=0 keeps every convolution on the fp32 CUDA-core kernels."""
    import os
    return os.environ.get("AMOE_TRAIN_TC", "1") != "0"


def _split_weight(weight: torch.Tensor, transposed: bool) -> torch.Tensor:
    """Six-term split packing of an nn.Conv2d weight.  Cached ON the tensor object (an attribute dies with it; a cache keyed
    by data_ptr would hand a new module the packed weights of a freed one) until its version changes: frozen experts pack once."""
    cache = getattr(weight, "_amoe_split6", None)
    if cache is None:
        cache = {}
        try:
            weight._amoe_split6 = cache
        except Exception:            # tensors that do not take attributes: pack every time
            pass
    hit = cache.get(bool(transposed))
    if hit is not None and hit[0] == (weight._version, weight.data_ptr()):
        return hit[1]
    Cout, Cin, KH, KW = weight.shape
    w = _f32c(weight.detach())
    rows, inner = (Cin, Cout) if transposed else (Cout, Cin)
    out = torch.empty((rows, KH, KW, 6, inner), device=w.device, dtype=torch.bfloat16)
    check(lib().amoe_pack_conv_weight_split6(ctx(w.device), ptr(w), ptr(out), Cout, Cin, KH, KW, int(transposed),
                                             stream_ptr(w.device)), "pack_conv_weight_split6")
    cache[bool(transposed)] = ((weight._version, weight.data_ptr()), out)
    return out


def _split3(x: torch.Tensor) -> torch.Tensor:
    """[..., C] fp32 -> [..., 3C] bf16 = (x1 | x2 | x3)."""
    C_ = x.shape[-1]
    rows = x.numel() // C_
    out = torch.empty(x.shape[:-1] + (3 * C_,), device=x.device, dtype=torch.bfloat16)
    check(lib().amoe_split3_bf16(ctx(x.device), ptr(x), ptr(out), rows, C_, stream_ptr(x.device)), "split3_bf16")
    return out


# ---- Cin = 3 first layers (ResNet stem 7x7/s2/p3, EasyBackbone conv1 5x5/s2/p2): fp32-accurate Toeplitz GEMM (csrc/stem_tc.cu) ----
def _split3_parts(t: torch.Tensor):
    """fp32 -> (t1, t2, t3) bf16 with t == t1 + t2 + t3 to 24 bits: each part the bf16 rounding of what is left (the differences
    are exact).  Operand preparation of the first layer: one padded frame per step, one filter image per weight version."""
    t1 = t.to(torch.bfloat16)
    r1 = t - t1.float()
    t2 = r1.to(torch.bfloat16)
    t3 = (r1 - t2.float()).to(torch.bfloat16)
    return t1, t2, t3


def image_nhwc4(image: torch.Tensor) -> torch.Tensor:
    """[B,3,H,W] NCHW -> [B,H,W,4] fp32 NHWC (zero 4th channel), cached on the image tensor object: the frozen experts and the
    policy backbone of one training step stage (and split, see _stem_split_frame) the same frame once."""
    from .. import _ops
    key = (image._version, image.data_ptr(), tuple(image.shape))
    cache = getattr(image, "_amoe_nhwc4", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    out = _ops.image_to_nhwc(image, 4, torch.float32)
    try:
        image._amoe_nhwc4 = (key, out)
    except Exception:
        pass
    return out


def _stem_tc_ok(Cx: int, Cin: int, Cout: int, KH: int, KW: int, stride: int, padding: int, H: int, W: int) -> bool:
    from .. import _ops
    if not train_tc() or Cx != 4 or Cin > 3 or stride != 2:
        return False
    if _ops.STEM_TOP - padding < 0 or _ops.STEM_LEFT - padding < 0 or KH - padding + _ops.STEM_TOP > _ops.STEM_KH or \
            KW - padding + _ops.STEM_LEFT > 8:
        return False
    return bool(lib().amoe_stem_fwd_f32tc_supported(H, W, _ops.STEM_KH, Cout))


def _stem_split_weight(weight: torch.Tensor, padding: int) -> torch.Tensor:
    """nn.Conv2d weight [Cout,Cin<=3,KH,KW] -> three filter images [3][STEM_KH*4][Cout][8] bf16 (layout of _ops.pack_stem: K index
    kh*32 + j*4 + c = padded row 2*oh+kh, padded pixel 2*ow+j, channel c).  Cached on the tensor until its version changes."""
    from .. import _ops
    key = (weight._version, weight.data_ptr(), padding)
    cache = getattr(weight, "_amoe_stem3", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    Cout, Cin, KH, KW = weight.shape
    w = _f32c(weight.detach())
    wk = w.new_zeros((Cout, _ops.STEM_KH, 8, 4))                                  # [n][kh][j][c]
    r0, j0 = _ops.STEM_TOP - padding, _ops.STEM_LEFT - padding
    wk[:, r0:r0 + KH, j0:j0 + KW, :Cin] = w.permute(0, 2, 3, 1)
    img = wk.reshape(Cout, _ops.STEM_KH * 4, 8).permute(1, 0, 2).contiguous()    # [k/8][n][8]
    out = torch.stack(_split3_parts(img), dim=0).contiguous()
    try:
        weight._amoe_stem3 = (key, out)
    except Exception:
        pass
    return out


def _stem_split_frame(x2: torch.Tensor) -> torch.Tensor:
    """[B,H,W,4] fp32 NHWC (4th channel zero) -> [3B,H+6,Wpad,4] bf16: the padded frame of the stem kernels (3 zero rows above and
    below, 4 zero pixels left) as its three bf16 parts stacked on the batch axis.  Cached on the tensor object: the three
    expert stems and the policy conv1 of one step read the same frame."""
    from .. import _ops
    key = (x2._version, x2.data_ptr())
    cache = getattr(x2, "_amoe_stem_frame3", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    B, H, W, Cx = x2.shape
    out = torch.empty((3 * B, H + 6, _ops.stem_wpad(W), Cx), device=x2.device, dtype=torch.bfloat16)
    check(lib().amoe_stem_split_frame(ctx(x2.device), ptr(x2), ptr(out), B, H, W, out.shape[2], stream_ptr(x2.device)),
          "stem_split_frame")
    try:
        x2._amoe_stem_frame3 = (key, out)
    except Exception:
        pass
    return out


def _even_size(H: int, W: int, KH: int, KW: int, stride: int, padding: int, Ho: int, Wo: int) -> Tuple[int, int]:
    """(H, W) rounded up to even for a stride-2 convolution when that leaves the output size unchanged (3x3/p1, 1x1/p0, ...):
    the extra zero row / column only stands where the zero padding already is."""
    if stride != 2:
        return H, W
    Hk, Wk = H + (H & 1), W + (W & 1)
    if (Hk + 2 * padding - KH) // 2 + 1 != Ho or (Wk + 2 * padding - KW) // 2 + 1 != Wo:
        return H, W
    return Hk, Wk


class _ConvBNAct(torch.autograd.Function):
    """nn.Conv2d(stride, padding, bias?) [-> nn.BatchNorm2d] [-> nn.ReLU] on NHWC fp32 activations.

    bn_mode: BN_NONE (plain convolution + bias), BN_BATCH (train-mode BatchNorm: statistics of this batch,
    running stats updated in place), BN_RUNNING (eval-mode BatchNorm, kept differentiable)."""

    @staticmethod
    def forward(ctx_, x, weight, bias, gamma, beta, running_mean, running_var, stride: int, padding: int,
                momentum: float, eps: float, bn_mode: int, relu: bool):
        x2 = _f32c(x)                                    # [B,H,W,Cx]  (Cx >= Cin: zero-padded channels)
        B, H, W, Cx = x2.shape
        Cout, Cin, KH, KW = weight.shape
        dev = x2.device
        h, st = ctx(dev), stream_ptr(dev)
        Ho, Wo = (H + 2 * padding - KH) // stride + 1, (W + 2 * padding - KW) // stride + 1
        ones = torch.ones(Cout, device=dev, dtype=torch.float32)
        cb = _f32c(bias) if bias is not None else torch.zeros(Cout, device=dev, dtype=torch.float32)
        conv = torch.empty((B, Ho, Wo, Cout), device=dev, dtype=torch.float32)
        fuse_relu = int(relu and bn_mode == BN_NONE)
        use_tc = train_tc() and Cx == Cin and bool(lib().amoe_conv2d_f32tc_supported(H, W, Cin, Cout, KH, KW, stride))
        # Stride 2 on an odd height / width (720p detection: 45 rows into layer4): the tensor-core kernels take even sizes, and
        # one appended zero row / column is exactly what the zero padding supplies there - same output, same gradient
        Hk, Wk = _even_size(H, W, KH, KW, stride, padding, Ho, Wo)
        if not use_tc and (Hk, Wk) != (H, W):
            use_tc = train_tc() and Cx == Cin and bool(lib().amoe_conv2d_f32tc_supported(Hk, Wk, Cin, Cout, KH, KW, stride))
        if not use_tc:
            Hk, Wk = H, W
        # (first layers whose weight is trained stay on the CUDA-core kernel: the gradient goldens of the policy backbone are
        # held to 5e-4 and move by 8e-4 with the - more accurate - tensor-core forward: ReLU units within rounding of zero)
        use_stem = (not use_tc) and (not ctx_.needs_input_grad[1]) and _stem_tc_ok(Cx, Cin, Cout, KH, KW, stride, padding, H, W)
        wp = None if (use_tc or use_stem) else _pack_w(weight, Cx)   # CUDA-core layout; packed lazily in backward on the tensor-core paths
        if use_stem:
            from .. import _ops
            xs, wsplit = _stem_split_frame(x2), _stem_split_weight(weight, padding)
            check(lib().amoe_stem_fwd_f32tc(h, ptr(xs), ptr(wsplit), ptr(ones), ptr(cb), ptr(conv), B, H, W, xs.shape[2],
                                            _ops.STEM_KH, Cout, fuse_relu, st), "stem_fwd_f32tc")
        elif use_tc:
            xin = x2 if (Hk, Wk) == (H, W) else torch.nn.functional.pad(x2, (0, 0, 0, Wk - W, 0, Hk - H))
            xs, wsplit = _split3(xin), _split_weight(weight, False)    # named: they must outlive the launch that reads them
            check(lib().amoe_conv2d_fwd_f32tc(h, ptr(xs), ptr(wsplit), ptr(ones), ptr(cb), ptr(conv),
                                              B, Hk, Wk, Cin, Cout, KH, KW, stride, padding, Ho, Wo, fuse_relu, st), "conv2d_fwd_f32tc")
            del xin
            del xs
        else:
            check(lib().amoe_conv2d_fwd(h, ptr(x2), ptr(wp), ptr(ones), ptr(cb), None, ptr(conv), 1, 0, B, H, W, Cx, Cout, KH, KW,
                                        stride, stride, padding, padding, Ho, Wo, fuse_relu, _cabi.F32, 1, 0, 0, st), "conv2d_fwd")
        ctx_.cfg = (stride, padding, bn_mode, relu, Cin, bias is not None)
        ctx_.weight_ref = weight if train_tc() else None
        ctx_.wshape = (Cout, KH, KW, Cx)
        if bn_mode == BN_NONE:
            ctx_.save_for_backward(x2, wp, conv if relu else None)
            return conv
        M = B * Ho * Wo
        g2, b2 = _f32c(gamma), _f32c(beta)
        y = torch.empty_like(conv)
        if bn_mode == BN_BATCH:
            ws = torch.empty(max(1, lib().amoe_colreduce_workspace_floats(M, Cout)), device=dev, dtype=torch.float32)
            mean = torch.empty(Cout, device=dev, dtype=torch.float32)
            rstd = torch.empty(Cout, device=dev, dtype=torch.float32)
            check(lib().amoe_bn_train_fwd(h, ptr(conv), ptr(g2), ptr(b2), ptr(running_mean), ptr(running_var),
                                          float(momentum), float(eps), ptr(y), ptr(mean), ptr(rstd), ptr(ws), M, Cout,
                                          int(relu), st), "bn_train_fwd")
        else:
            mean = _f32c(running_mean)
            rstd = torch.rsqrt(_f32c(running_var) + eps)   # C values: parameter preparation, not activation math
            check(lib().amoe_bn_apply_fwd(h, ptr(conv), ptr(mean), ptr(rstd), ptr(g2), ptr(b2), ptr(y), M, Cout, int(relu), st),
                  "bn_apply_fwd")
        ctx_.save_for_backward(x2, wp, conv, y if relu else None, g2, mean, rstd)
        return y

    @staticmethod
    def backward(ctx_, dy):
        stride, padding, bn_mode, relu, Cin, has_bias = ctx_.cfg
        dy = _f32c(dy)
        dgamma = dbeta = None
        if bn_mode == BN_NONE:
            x2, wp, y = ctx_.saved_tensors
            dev = x2.device
            h, st = ctx(dev), stream_ptr(dev)
            if relu:
                dconv = torch.empty_like(dy)
                check(lib().amoe_relu_bwd(h, ptr(dy), ptr(y), ptr(dconv), dy.numel(), st), "relu_bwd")
            else:
                dconv = dy
            B, Ho, Wo, Cout = dconv.shape
            M = B * Ho * Wo
            ws = torch.empty(max(1, lib().amoe_colreduce_workspace_floats(M, Cout)), device=dev, dtype=torch.float32)
        else:
            x2, wp, conv, y, g2, mean, rstd = ctx_.saved_tensors
            dev = x2.device
            h, st = ctx(dev), stream_ptr(dev)
            B, Ho, Wo, Cout = conv.shape
            M = B * Ho * Wo
            ws = torch.empty(max(1, lib().amoe_colreduce_workspace_floats(M, Cout)), device=dev, dtype=torch.float32)
            dconv = torch.empty_like(conv)
            dgamma = torch.empty(Cout, device=dev, dtype=torch.float32)
            dbeta = torch.empty(Cout, device=dev, dtype=torch.float32)
            check(lib().amoe_bn_bwd(h, ptr(dy), ptr(conv), ptr(y), ptr(g2), ptr(mean), ptr(rstd), ptr(dconv), ptr(dgamma),
                                    ptr(dbeta), ptr(ws), M, Cout, int(bn_mode == BN_BATCH), st), "bn_bwd")
        B, H, W, Cx = x2.shape
        KH, KW = ctx_.wshape[1], ctx_.wshape[2]
        w_ref = getattr(ctx_, "weight_ref", None)

        def packed_w():
            return wp if wp is not None else _pack_w(w_ref, Cx)
        dbias = None
        if has_bias and ctx_.needs_input_grad[2]:
            dbias = torch.empty(Cout, device=dev, dtype=torch.float32)
            check(lib().amoe_colsum(h, ptr(dconv), ptr(dbias), ptr(ws), M, Cout, 1.0, st), "colsum")
        dw = None
        wg_tc = (ctx_.needs_input_grad[1] and w_ref is not None and KH == 3 and KW == 3 and stride == 1 and padding == 1 and
                 Cx == Cin and bool(lib().amoe_conv3x3_wgrad_f32tc_supported(Cx, Cout)))
        if wg_tc:
            # weight gradient on the tensor cores: a GEMM per tap over the padded position grid (csrc/wgrad_tc.cu)
            P = B * (H + 2) * (W + 2)
            x3 = torch.empty((B, H + 2, W + 2, 3 * Cx), device=dev, dtype=torch.bfloat16)
            dy3 = torch.empty((B, H + 2, W + 2, 3 * Cout), device=dev, dtype=torch.bfloat16)
            check(lib().amoe_split3_padded(h, ptr(x2), ptr(x3), B, H, W, Cx, st), "split3_padded(x)")
            check(lib().amoe_split3_padded(h, ptr(dconv), ptr(dy3), B, Ho, Wo, Cout, st), "split3_padded(dy)")
            n_ws = int(lib().amoe_conv3x3_wgrad_f32tc_workspace_floats(h, Cx, Cout, P))
            ws2 = torch.empty(n_ws, device=dev, dtype=torch.float32)
            dwp = torch.empty(ctx_.wshape, device=dev, dtype=torch.float32)
            check(lib().amoe_conv3x3_wgrad_f32tc(h, ptr(dy3), ptr(x3), ptr(dwp), ptr(ws2), n_ws, W, Cx, Cout, P, st),
                  "conv3x3_wgrad_f32tc")
            dw = dwp.permute(0, 3, 1, 2).contiguous()
            del x3, dy3
        elif ctx_.needs_input_grad[1]:
            n_ws = lib().amoe_conv2d_bwd_weight_workspace_floats(h, B, Cx, Cout, KH, KW, Ho, Wo)
            ws2 = torch.empty(max(1, n_ws), device=dev, dtype=torch.float32)
            dwp = torch.empty(ctx_.wshape, device=dev, dtype=torch.float32)
            check(lib().amoe_conv2d_bwd_weight(h, ptr(dconv), ptr(x2), ptr(dwp), ptr(ws2), n_ws, B, H, W, Cx, Cout, KH, KW,
                                               stride, stride, padding, padding, Ho, Wo, st), "conv2d_bwd_weight")
            dw = dwp[..., :Cin].permute(0, 3, 1, 2).contiguous()   # packed [Cout,KH,KW,Cin] -> OIHW
        dx = None
        if ctx_.needs_input_grad[0]:
            # dgrad on the tensor cores: roles swap (input = dy with Cout channels, output channels = Cin)
            Hk, Wk = _even_size(H, W, KH, KW, stride, padding, Ho, Wo)          # odd sizes: one zero row / column appended
            tc_ok = (w_ref is not None and Cout % 64 == 0 and Cx % 32 == 0 and (Cx <= 256 or Cx % 256 == 0) and
                     KH * KW * 6 <= 64 and (stride == 1 or (stride == 2 and Hk % 2 == 0 and Wk % 2 == 0)))
            if tc_ok:
                holes = stride == 2 and (KH < 2 or KW < 2)         # 1x1 / stride 2: three of four parity classes get no tap
                dx = (torch.zeros if holes else torch.empty)((B, Hk, Wk, Cx), device=dev, dtype=torch.float32)
                ones_i = torch.ones(Cx, device=dev, dtype=torch.float32)
                zeros_i = torch.zeros(Cx, device=dev, dtype=torch.float32)
                dys, wts = _split3(dconv), _split_weight(w_ref, True)  # named: they must outlive the launches that read them
                check(lib().amoe_conv2d_bwd_data_f32tc(h, ptr(dys), ptr(wts), ptr(ones_i), ptr(zeros_i), ptr(dx), B, Hk, Wk, Cx, Cout,
                                                       KH, KW, stride, padding, Ho, Wo, st), "conv2d_bwd_data_f32tc")
                del dys
                if (Hk, Wk) != (H, W):
                    dx = dx[:, :H, :W, :].contiguous()              # the appended row / column is not an input: its gradient is dropped
            else:
                dx = torch.empty_like(x2)
                wpk = packed_w()
                check(lib().amoe_conv2d_bwd_data(h, ptr(dconv), ptr(wpk), ptr(dx), B, H, W, Cx, Cout, KH, KW, stride, stride,
                                                 padding, padding, Ho, Wo, st), "conv2d_bwd_data")
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None


def conv_bn_act(x_nhwc, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], relu: bool, batch_stats: Optional[bool] = None):
    """conv [+ BatchNorm] [+ ReLU].  batch_stats defaults to bn.training."""
    if conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1]:
        raise NotImplementedError("training conv kernels take square stride/padding")
    if bn is None:
        return _ConvBNAct.apply(x_nhwc, conv.weight, conv.bias, None, None, None, None, conv.stride[0], conv.padding[0],
                                0.0, 0.0, BN_NONE, relu)
    if batch_stats is None:
        batch_stats = bn.training
    momentum = 0.1 if bn.momentum is None else bn.momentum
    rm, rv = bn.running_mean, bn.running_var
    if rm is None or rv is None:
        raise NotImplementedError("BatchNorm2d(track_running_stats=False) is not supported by the training kernels")
    if batch_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    y = _ConvBNAct.apply(x_nhwc, conv.weight, conv.bias, bn.weight, bn.bias, rm, rv, conv.stride[0], conv.padding[0],
                         momentum, bn.eps, BN_BATCH if batch_stats else BN_RUNNING, relu)
    if batch_stats:   # the kernel updated the running statistics through raw pointers: bump Tensor._version
        torch.autograd.graph.increment_version([rm, rv])
    return y


def conv_bn_relu(x_nhwc, conv: nn.Conv2d, bn: nn.BatchNorm2d, batch_stats: bool) -> torch.Tensor:
    return conv_bn_act(x_nhwc, conv, bn, relu=True, batch_stats=batch_stats)


class _MaxPool3x3s2(torch.autograd.Function):
    """nn.MaxPool2d(3, stride 2, padding 1) on NHWC fp32."""

    @staticmethod
    def forward(ctx_, x):
        x2 = _f32c(x)
        B, H, W, Cc = x2.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((B, Ho, Wo, Cc), device=x2.device, dtype=torch.float32)
        check(lib().amoe_maxpool3x3s2_fwd(ctx(x2.device), ptr(x2), ptr(y), B, H, W, Cc, _cabi.F32, 0, stream_ptr(x2.device)),
              "maxpool3x3s2_fwd")
        ctx_.save_for_backward(x2)
        return y

    @staticmethod
    def backward(ctx_, dy):
        (x2,) = ctx_.saved_tensors
        dy = _f32c(dy)
        B, H, W, Cc = x2.shape
        dx = torch.empty_like(x2)
        ws = torch.empty(dy.shape, device=x2.device, dtype=torch.uint8)      # arg-max tap of every pooling window
        check(lib().amoe_maxpool3x3s2_bwd_ws(ctx(x2.device), ptr(x2), ptr(dy), ptr(dx), ptr(ws), B, H, W, Cc, stream_ptr(x2.device)),
              "maxpool3x3s2_bwd_ws")
        return dx


def max_pool3x3s2(x_nhwc) -> torch.Tensor:
    return _MaxPool3x3s2.apply(x_nhwc)


class _AddReLU(torch.autograd.Function):
    """BasicBlock tail: relu(main + identity)."""

    @staticmethod
    def forward(ctx_, a, b):
        a2, b2 = _f32c(a), _f32c(b)
        y = torch.empty_like(a2)
        check(lib().amoe_add_relu_fwd(ctx(a2.device), ptr(a2), ptr(b2), ptr(y), a2.numel(), stream_ptr(a2.device)), "add_relu_fwd")
        ctx_.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx_, dy):
        (y,) = ctx_.saved_tensors
        dy = _f32c(dy)
        g = torch.empty_like(y)
        check(lib().amoe_relu_bwd(ctx(y.device), ptr(dy), ptr(y), ptr(g), y.numel(), stream_ptr(y.device)), "relu_bwd")
        return g, g


def add_relu(a, b) -> torch.Tensor:
    return _AddReLU.apply(a, b)


class _UpsampleBilinearNCHW(torch.autograd.Function):
    """F.interpolate(low, size=(H,W), mode="bilinear", align_corners=False) from the NHWC head output to the
    NCHW logits the segmentation / drivable experts return."""

    @staticmethod
    def forward(ctx_, low, H: int, W: int):
        lo = _f32c(low)
        B, h, w, Cc = lo.shape
        out = torch.empty((B, Cc, H, W), device=lo.device, dtype=torch.float32)
        check(lib().amoe_upsample_bilinear_nchw_fwd(ctx(lo.device), ptr(lo), ptr(out), B, h, w, Cc, H, W, _cabi.F32,
                                                    stream_ptr(lo.device)), "upsample_bilinear_nchw_fwd")
        ctx_.shape = (B, h, w, Cc, H, W)
        return out

    @staticmethod
    def backward(ctx_, dy):
        B, h, w, Cc, H, W = ctx_.shape
        dy = _f32c(dy)
        dlow = torch.empty((B, h, w, Cc), device=dy.device, dtype=torch.float32)
        check(lib().amoe_upsample_bilinear_nchw_bwd(ctx(dy.device), ptr(dy), ptr(dlow), B, h, w, Cc, H, W, stream_ptr(dy.device)),
              "upsample_bilinear_nchw_bwd")
        return dlow, None, None


def upsample_bilinear_nchw(low_nhwc, H: int, W: int) -> torch.Tensor:
    return _UpsampleBilinearNCHW.apply(low_nhwc, H, W)


class _DetLoss(torch.autograd.Function):
    """CrossEntropy(ignore_index=num_classes) + w * SmoothL1 over the matched queries; `head_out` is the
    NHWC head output [B,h,w,C+4] (class logits then box deltas in the channel axis)."""

    @staticmethod
    def forward(ctx_, head_out, target_classes, target_boxes, num_classes: int, bbox_weight: float):
        ho = _f32c(head_out)
        B, h, w, Ct = ho.shape
        rows = B * h * w
        dev = ho.device
        losses = torch.empty(4, device=dev, dtype=torch.float32)
        dho = torch.empty_like(ho)
        base = ho.data_ptr()
        check(lib().amoe_det_loss_fwd_bwd(ctx(dev), base, Ct, base + 4 * num_classes, Ct, ptr(target_classes), ptr(target_boxes),
                                          rows, num_classes, num_classes, float(bbox_weight), ptr(losses), dho.data_ptr(), Ct,
                                          dho.data_ptr() + 4 * num_classes, Ct, stream_ptr(dev)), "det_loss_fwd_bwd")
        ctx_.save_for_backward(dho)
        return losses

    @staticmethod
    def backward(ctx_, dlosses):
        (dho,) = ctx_.saved_tensors
        return dho * dlosses[0], None, None, None, None


class _GlobalAvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx_, x):
        x2 = _f32c(x)
        B, H, W, Cc = x2.shape
        out = torch.empty((B, Cc), device=x2.device, dtype=torch.float32)
        check(lib().amoe_gap_fwd(ctx(x2.device), ptr(x2), ptr(out), B, H * W, Cc, stream_ptr(x2.device)), "gap_fwd")
        ctx_.shape = (B, H, W, Cc)
        return out

    @staticmethod
    def backward(ctx_, dy):
        B, H, W, Cc = ctx_.shape
        dy = _f32c(dy)
        dx = torch.empty((B, H, W, Cc), device=dy.device, dtype=torch.float32)
        check(lib().amoe_gap_bwd(ctx(dy.device), ptr(dy), ptr(dx), B, H * W, Cc, stream_ptr(dy.device)), "gap_bwd")
        return dx


def global_avg_pool(x_nhwc) -> torch.Tensor:
    return _GlobalAvgPool.apply(x_nhwc)


# ------------------------------------------------------------------------------------------------
LOSS_NAMES = ("total_loss", "ade", "fde", "speed", "smoothness", "load_balancing", "entropy")


class _GatingLoss(torch.autograd.Function):
    """All seven terms of compute_gating_losses in one launch; the gradient of total_loss w.r.t. the
    predictions is produced by the same launch and handed to autograd in backward."""

    @staticmethod
    def forward(ctx_, waypoints, speed, expert_weights, tgt_wp, tgt_spd, speed_mode: int, coef, use_lb: bool, use_ent: bool):
        wp, ew, twp = _f32c(waypoints), _f32c(expert_weights), _f32c(tgt_wp)
        B, H, _ = wp.shape
        E = ew.shape[1]
        dev = wp.device
        spd = _f32c(speed) if speed is not None else None
        tspd = _f32c(tgt_spd) if tgt_spd is not None else None
        losses = torch.empty(7, device=dev, dtype=torch.float32)
        dwp = torch.empty_like(wp)
        dspd = torch.zeros((B, spd.shape[1]), device=dev, dtype=torch.float32) if spd is not None else None
        dew = torch.empty_like(ew)
        if spd is not None and spd.shape[1] != H and speed_mode == 1:
            raise ValueError("speed sequence length differs from the waypoint horizon")
        coef_c = (C.c_float * 6)(*[float(c) for c in coef])
        Hs = spd.shape[1] if spd is not None else 0
        d_spd_arg = dspd if (spd is not None and speed_mode in (1, 2)) else None
        check(lib().amoe_gating_loss_fwd_bwd(ctx(dev), ptr(wp), ptr(spd), Hs, ptr(ew), ptr(twp), ptr(tspd),
                                             tspd.shape[1] if tspd is not None else 0, B, H, E, int(speed_mode), coef_c,
                                             int(use_lb), int(use_ent), ptr(losses), ptr(dwp), ptr(d_spd_arg), ptr(dew),
                                             stream_ptr(dev)), "gating_loss_fwd_bwd")
        ctx_.save_for_backward(dwp, dspd, dew)
        ctx_.mark_non_differentiable()
        return losses

    @staticmethod
    def backward(ctx_, dlosses):
        dwp, dspd, dew = ctx_.saved_tensors
        # the reference back-propagates total_loss only; gradients arriving at the other six entries (a caller
        # differentiating e.g. `ade` alone) are not supported by the fused kernel
        s = dlosses[0]
        return dwp * s, (dspd * s if dspd is not None else None), dew * s, None, None, None, None, None, None
