"""GPU (-m gpu): each sm_100a kernel, called through the C-ABI, against stock torch ops /
the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from _util import rel_err, rel_l2

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _conv_case(G, B, H, W, Cin, Cout, k, s, p, dtype, residual, relu, bn=True, bias=False, impl=0, seed=0):
    """Build G convs (+BN) with random weights, run ours (grouped) and torch (per group)."""
    from automoe_b200 import _ops
    g = torch.Generator().manual_seed(seed)
    convs, bns = [], []
    for _ in range(G):
        c = nn.Conv2d(Cin, Cout, k, s, p, bias=bias)
        with torch.no_grad():
            c.weight.copy_(torch.randn(c.weight.shape, generator=g) * (2.0 / (Cin * k * k)) ** 0.5)
            if bias:
                c.bias.copy_(torch.randn(Cout, generator=g) * 0.1)
        convs.append(c.to(DEV))
        if bn:
            b = nn.BatchNorm2d(Cout)
            with torch.no_grad():
                b.weight.copy_(1 + 0.1 * torch.randn(Cout, generator=g))
                b.bias.copy_(0.1 * torch.randn(Cout, generator=g))
                b.running_mean.copy_(0.1 * torch.randn(Cout, generator=g))
                b.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
            bns.append(b.to(DEV).eval())
    x = torch.randn((G * B, Cin, H, W), generator=g).to(DEV)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = torch.randn((G * B, Cout, Ho, Wo), generator=g).to(DEV) if residual else None
    if dtype == torch.bfloat16:  # both sides see the same bf16-rounded operands
        x = x.bfloat16().float()
        if res is not None:
            res = res.bfloat16().float()
        for c in convs:
            c.weight.data = c.weight.data.bfloat16().float()
    # ours
    cin_pad = 4 if Cin == 3 else None
    pc = _ops.pack_conv(convs, bns if bn else None, dtype, torch.device(DEV), relu=relu, cin_pad=cin_pad)
    xn = x.permute(0, 2, 3, 1).contiguous()
    if cin_pad:
        xn = F.pad(xn, (0, cin_pad - Cin))
    xn = xn.to(dtype).contiguous()
    rn = res.permute(0, 2, 3, 1).contiguous().to(dtype) if res is not None else None
    y = _ops.conv2d(pc, xn, B, H, W, residual=rn, impl=impl)
    torch.cuda.synchronize()
    # reference
    refs = []
    with torch.no_grad():
        for gi in range(G):
            r = convs[gi](x[gi * B:(gi + 1) * B])
            if bn:
                r = bns[gi](r)
            if res is not None:
                r = r + res[gi * B:(gi + 1) * B]
            if relu:
                r = F.relu(r)
            refs.append(r)
    ref = torch.cat(refs, 0)
    return y.float().permute(0, 3, 1, 2), ref


SIMT_SHAPES = [
    # G, B, H, W, Cin, Cout, k, s, p, residual, relu, bias
    (1, 2, 32, 32, 3, 64, 7, 2, 3, False, True, False),    # stem
    (3, 2, 32, 32, 3, 64, 7, 2, 3, False, True, False),    # grouped stem
    (1, 2, 16, 16, 64, 64, 3, 1, 1, True, True, False),    # BasicBlock conv2 + residual
    (1, 3, 16, 16, 64, 128, 3, 2, 1, False, True, False),
    (1, 3, 16, 16, 64, 128, 1, 2, 0, False, False, False),  # downsample
    (1, 2, 32, 32, 3, 32, 5, 2, 2, False, True, True),     # policy conv1 (bias + BN)
    (2, 1, 7, 9, 20, 24, 3, 1, 1, False, False, True),     # odd everything
]


@pytest.mark.parametrize("shape", SIMT_SHAPES)
def test_conv_simt_fp32(shape):
    G, B, H, W, Cin, Cout, k, s, p, residual, relu, bias = shape
    y, ref = _conv_case(G, B, H, W, Cin, Cout, k, s, p, torch.float32, residual, relu, bias=bias)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 2e-5, rel_err(y, ref)


@pytest.mark.parametrize("shape", SIMT_SHAPES[:2] + SIMT_SHAPES[5:6])
def test_conv_simt_bf16(shape):
    G, B, H, W, Cin, Cout, k, s, p, residual, relu, bias = shape
    y, ref = _conv_case(G, B, H, W, Cin, Cout, k, s, p, torch.bfloat16, residual, relu, bias=bias, impl=1)
    assert rel_err(y, ref) < 8e-3, rel_err(y, ref)


TC_SHAPES = [
    # G, B, H, W, Cin, Cout, k, s, p, residual, relu
    (1, 2, 16, 16, 64, 64, 3, 1, 1, False, True),      # smallest: 1 k-chunk, N=64
    (1, 2, 16, 16, 64, 64, 3, 1, 1, True, True),       # + residual
    (1, 2, 16, 16, 64, 64, 1, 1, 0, False, False),     # 1x1: one tap
    (1, 4, 64, 64, 64, 64, 3, 1, 1, True, True),       # layer1 geometry (tw=64, th=2)
    (1, 2, 32, 32, 64, 128, 3, 2, 1, False, True),     # layer2.0.conv1: stride-2 parity view
    (1, 2, 32, 32, 64, 128, 1, 2, 0, False, False),    # layer2.0.downsample
    (1, 2, 16, 16, 128, 128, 3, 1, 1, True, True),     # 2 k-chunks, N=128
    (1, 3, 16, 16, 128, 256, 3, 2, 1, False, True),    # N=256
    (1, 3, 8, 8, 256, 512, 3, 2, 1, False, True),      # 2 N-tiles of 256
    (3, 2, 8, 8, 512, 512, 3, 1, 1, True, True),       # grouped layer4 (nb=2)
    (3, 1, 8, 8, 512, 256, 3, 1, 1, False, True),      # expert head conv (bias, no BN) grouped, B=1
    (1, 5, 14, 14, 64, 64, 3, 1, 1, True, True),       # 224-px geometry, overhanging tiles, odd batch
    (1, 2, 7, 7, 128, 128, 3, 1, 1, False, True),
    (1, 1, 56, 56, 64, 64, 3, 1, 1, False, True),
    (2, 2, 4, 4, 256, 192, 3, 1, 1, False, False),     # N=192
    (1, 2, 2, 2, 512, 512, 3, 1, 1, True, True),       # 64-px input at layer4: 2x2 map, nb=32
    (1, 2, 32, 32, 32, 64, 3, 2, 1, False, True),      # policy conv2 through the pixel-pair view
    (1, 2, 16, 16, 64, 128, 3, 2, 1, False, True),     # policy conv3
    (1, 40, 16, 16, 64, 64, 3, 1, 1, False, True),     # > 148 tiles: persistent loop, both accumulators
]


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tc_bf16(shape):
    G, B, H, W, Cin, Cout, k, s, p, residual, relu = shape
    is_head = (Cout == 256 and Cin == 512)
    y, ref = _conv_case(G, B, H, W, Cin, Cout, k, s, p, torch.bfloat16, residual, relu, bn=not is_head,
                        bias=is_head or Cin == 32, impl=2)
    assert y.shape == ref.shape
    e, l2 = rel_err(y, ref), rel_l2(y, ref)
    assert torch.isfinite(y).all()
    # fp32 accumulation of identical bf16 operands; only the output rounding (2^-9) and summation order differ
    assert e < 8e-3 and l2 < 4e-3, (e, l2)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", [
    # G, B, H, W, Cin, Cout, k, s
    (3, 64, 8, 8, 512, 512, 3, 1),      # 192 pair-tiles of N=256 on 74 CTA pairs: partial last wave -> tail splitting
    (3, 40, 16, 16, 256, 256, 3, 1),    # CTA pairs, one N tile
    (2, 40, 32, 32, 64, 128, 3, 2),     # stage entry geometry: resident weights, grid (CTAs, G)
    (1, 300, 8, 8, 128, 64, 1, 1),      # 1x1
])
def test_conv_tc_walk_direction_and_schedules(case, monkeypatch):
    """The tile schedule of csrc/conv_tc.cu (walk direction, CTA pairs, tail splitting, resident weights) never changes a
    result bit: every variant against the plain one-CTA-per-tile front-to-back launch."""
    from automoe_b200 import _ops
    G, B, H, W, Cin, Cout, k, s = case
    g = torch.Generator().manual_seed(21)
    convs, bns = _mk_conv_bn(Cin, Cout, k, s, k // 2, g, n=G)
    x = torch.randn((G * B, H, W, Cin), generator=g).to(DEV).bfloat16()
    pc = _ops.pack_conv(convs, bns, torch.bfloat16, torch.device(DEV), relu=True)
    for key in ("AMOE_TC_PAIR", "AMOE_TC_TAIL_SPLIT", "AMOE_TC_WRES"):
        monkeypatch.setenv(key, "0")
    _ops.walk_reset(DEV, 0)
    ref = _ops.conv2d(pc, x, B, H, W)
    for key in ("AMOE_TC_PAIR", "AMOE_TC_TAIL_SPLIT", "AMOE_TC_WRES"):
        monkeypatch.delenv(key)
    for first in (0, 1):
        _ops.walk_reset(DEV, first)
        y = _ops.conv2d(pc, x, B, H, W)
        torch.cuda.synchronize()
        assert torch.equal(y, ref), (first, (y.float() - ref.float()).abs().max().item())
    monkeypatch.setenv("AMOE_TC_PAIR", "1")          # pairs wherever the kernel takes them
    _ops.walk_reset(DEV, 1)
    y = _ops.conv2d(pc, x, B, H, W)
    torch.cuda.synchronize()
    assert torch.equal(y, ref)
    with torch.no_grad():
        want = torch.cat([F.relu(bns[i](convs[i](x[i * B:(i + 1) * B].float().permute(0, 3, 1, 2)))) for i in range(G)], 0)
    assert rel_err(y.float().permute(0, 3, 1, 2), want) < 8e-3


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", [
    # G, B, H, W, Cin, Cout, k, s, residual
    (3, 6, 8, 8, 512, 512, 3, 1, True),       # layer4 block conv2: N = 256 tiles, CTA pairs, bf16 residual
    (2, 9, 32, 32, 64, 128, 3, 2, False),     # stage entry geometry: resident weights, N = 128
    (1, 7, 16, 16, 128, 64, 1, 1, True),      # 1x1, N = 64 (one 64-column pass of the chunk loop)
    (2, 2, 4, 4, 256, 192, 3, 1, False),      # N = 192: an odd number of 32-column chunks
])
def test_conv_tc_lean_epilogue_is_bit_identical(case, monkeypatch):
    """The bf16 inference epilogue of csrc/conv_tc.cu (two-lane FMAs, ReLU on packed pairs, 32-byte stores and residual
    loads) writes the bits of the generic epilogue (scalar FFMA / FMNMX, 16-byte accesses)."""
    from automoe_b200 import _ops
    G, B, H, W, Cin, Cout, k, s, residual = case
    g = torch.Generator().manual_seed(33)
    convs, bns = _mk_conv_bn(Cin, Cout, k, s, k // 2, g, n=G)
    x = torch.randn((G * B, H, W, Cin), generator=g).to(DEV).bfloat16()
    Ho, Wo = H // s, W // s
    res = torch.randn((G * B, Ho, Wo, Cout), generator=g).to(DEV).bfloat16() if residual else None
    pc = _ops.pack_conv(convs, bns, torch.bfloat16, torch.device(DEV), relu=True)
    monkeypatch.setenv("AMOE_TC_LEAN", "0")
    _ops.walk_reset(DEV, 0)
    ref = _ops.conv2d(pc, x, B, H, W, residual=res)
    for lean, w32 in (("1", "0"), ("1", "1")):
        monkeypatch.setenv("AMOE_TC_LEAN", lean)
        monkeypatch.setenv("AMOE_TC_W32", w32)
        _ops.walk_reset(DEV, 0)
        y = _ops.conv2d(pc, x, B, H, W, residual=res)
        torch.cuda.synchronize()
        assert torch.equal(y, ref), (lean, w32, (y.float() - ref.float()).abs().max().item())
    assert (ref.float() < 0).sum() == 0        # ReLU on the packed pairs never leaves a negative (or -0 -> sign bit) value
    assert (ref.view(torch.int16) < 0).sum() == 0


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", [("stem", 3, 2, 64, 64), ("stem", 3, 3, 256, 256), ("stem", 1, 2, 224, 224),
                                  ("policy", 1, 2, 64, 64), ("policy", 1, 3, 256, 256)])
def test_conv_tc_rowwin(case):
    """Cin=3 convolutions (ResNet stem 7x7/s2/p3, policy conv1 5x5/s2/p2) on the tensor cores through
    row windows of the physically padded image; several convs sharing the input are one GEMM."""
    from automoe_b200 import _ops
    kind, n, B, H, W = case
    g = torch.Generator().manual_seed(4)
    if kind == "stem":
        mk = lambda: nn.Conv2d(3, 64, 7, 2, 3, bias=False)
    else:
        mk = lambda: nn.Conv2d(3, 32, 5, 2, 2, bias=True)
    convs, bns = [], []
    for _ in range(n):
        c = mk()
        cout = c.weight.shape[0]
        with torch.no_grad():
            c.weight.copy_((torch.randn(c.weight.shape, generator=g) * 0.1).bfloat16().float())
            if c.bias is not None:
                c.bias.copy_(torch.randn(cout, generator=g) * 0.1)
        b = nn.BatchNorm2d(cout)
        with torch.no_grad():
            b.weight.copy_(1 + 0.1 * torch.randn(cout, generator=g))
            b.bias.copy_(0.1 * torch.randn(cout, generator=g))
            b.running_mean.copy_(0.1 * torch.randn(cout, generator=g))
            b.running_var.copy_(torch.rand(cout, generator=g) + 0.5)
        convs.append(c.to(DEV))
        bns.append(b.to(DEV).eval())
    img = torch.randn((B, 3, H, W), generator=g).bfloat16().float().to(DEV)
    pc = _ops.pack_rowwin(convs, bns, torch.device(DEV), relu=True)
    xp = _ops.image_to_nhwc_padded(img, 4, _ops.ROWWIN_LEFT, _ops.rowwin_wpad(W), torch.bfloat16)
    assert torch.equal(xp[:, :, 4:4 + W, :3].float(), img.permute(0, 2, 3, 1)) and (xp[:, :, :4] == 0).all() and (xp[:, :, 4 + W:] == 0).all()
    y = _ops.conv2d_rowwin(pc, xp, B, H, W).float().permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = torch.cat([F.relu(bn(c(img))) for c, bn in zip(convs, bns)], 0)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 8e-3 and rel_l2(y, ref) < 4e-3, (rel_err(y, ref), rel_l2(y, ref))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", [(3, True, 2, 64, 64), (3, True, 3, 256, 256), (1, False, 2, 224, 224), (0, True, 2, 64, 96),
                                  (3, True, 1, 32, 32)])
def test_stem_tc(case):
    """All Cin=3 first-layer convolutions (expert stems 7x7/s2/p3 + policy conv1 5x5/s2/p2) as one GEMM over
    the raw image rows (csrc/stem_tc.cu, Toeplitz no-swizzle A operand)."""
    from automoe_b200 import _ops
    n_exp, with_policy, B, H, W = case
    g = torch.Generator().manual_seed(10)
    convs, bns = [], []
    for i in range(n_exp + int(with_policy)):
        c = nn.Conv2d(3, 64, 7, 2, 3, bias=False) if i < n_exp else nn.Conv2d(3, 32, 5, 2, 2, bias=True)
        cout = c.weight.shape[0]
        b = nn.BatchNorm2d(cout)
        with torch.no_grad():
            c.weight.copy_((torch.randn(c.weight.shape, generator=g) * 0.1).bfloat16().float())
            if c.bias is not None:
                c.bias.copy_(torch.randn(cout, generator=g) * 0.1)
            b.weight.copy_(1 + 0.1 * torch.randn(cout, generator=g))
            b.bias.copy_(0.1 * torch.randn(cout, generator=g))
            b.running_mean.copy_(0.1 * torch.randn(cout, generator=g))
            b.running_var.copy_(torch.rand(cout, generator=g) + 0.5)
        convs.append(c.to(DEV))
        bns.append(b.to(DEV).eval())
    img = torch.randn((B, 3, H, W), generator=g).bfloat16().float().to(DEV)
    ps = _ops.pack_stem(convs, bns, torch.device(DEV), relu=True)
    xp = _ops.stage_image_stem(img)
    assert torch.equal(xp[:, 3:3 + H, 4:4 + W, :3].float(), img.permute(0, 2, 3, 1))
    rgb = xp[..., :3]   # the image channels have a zero border; the padding channel is 1.0 everywhere (folded-bias carrier)
    assert (rgb[:, :3] == 0).all() and (rgb[:, 3 + H:] == 0).all() and (rgb[:, :, :4] == 0).all() and (rgb[:, :, 4 + W:] == 0).all()
    assert (xp[..., 3] == 1).all()
    groups = ([n_exp] if n_exp else []) + ([1] if with_policy else [])
    outs = _ops.stem_forward(ps, xp, B, H, W, groups=groups)
    torch.cuda.synchronize()
    with torch.no_grad():
        refs = [F.relu(bn(c(img))) for c, bn in zip(convs, bns)]
    if n_exp:
        ref = torch.cat(refs[:n_exp], 0)
        y = outs[0].float().permute(0, 3, 1, 2)
        assert y.shape == ref.shape
        assert rel_err(y, ref) < 8e-3 and rel_l2(y, ref) < 4e-3, (rel_err(y, ref), rel_l2(y, ref))
    if with_policy:
        y = outs[-1].float().permute(0, 3, 1, 2)
        assert y.shape == refs[-1].shape
        assert rel_err(y, refs[-1]) < 8e-3 and rel_l2(y, refs[-1]) < 4e-3, (rel_err(y, refs[-1]), rel_l2(y, refs[-1]))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", [(3, True, 2, 64, 64, 1), (3, True, 3, 256, 256, 1), (1, False, 2, 224, 224, 0),
                                  (3, True, 5, 32, 32, 1), (1, True, 150, 16, 16, 0), (3, True, 2, 256, 128, 1)])
@pytest.mark.parametrize("fold", ["1", "0"])
def test_stem_pool_fused(case, fold, monkeypatch):
    """fold=1: BatchNorm scale folded into the bf16 filters, bias carried by the frame's padding channel;
    fold=0: fp32 scale/bias epilogue.  Stem GEMM with the ResNet max-pool fused behind the expert channels (policy conv1 stays full-res)."""
    from automoe_b200 import _ops
    monkeypatch.setenv("AMOE_STEM_FOLD", fold)
    n_exp, with_policy, B, H, W, out_pad = case
    g = torch.Generator().manual_seed(12)
    convs, bns = [], []
    for i in range(n_exp + int(with_policy)):
        c = nn.Conv2d(3, 64, 7, 2, 3, bias=False) if i < n_exp else nn.Conv2d(3, 32, 5, 2, 2, bias=True)
        cout = c.weight.shape[0]
        b = nn.BatchNorm2d(cout)
        with torch.no_grad():
            c.weight.copy_((torch.randn(c.weight.shape, generator=g) * 0.1).bfloat16().float())
            if c.bias is not None:
                c.bias.copy_(torch.randn(cout, generator=g) * 0.1)
            b.weight.copy_(1 + 0.1 * torch.randn(cout, generator=g))
            b.bias.copy_(0.1 * torch.randn(cout, generator=g))
            b.running_mean.copy_(0.1 * torch.randn(cout, generator=g))
            b.running_var.copy_(torch.rand(cout, generator=g) + 0.5)
        convs.append(c.to(DEV))
        bns.append(b.to(DEV).eval())
    img = torch.randn((B, 3, H, W), generator=g).bfloat16().float().to(DEV)
    ps = _ops.pack_stem(convs, bns, torch.device(DEV), relu=True)
    xp = _ops.stage_image_stem(img)
    assert _ops.stem_pool_supported(H, W)
    pooled, rest = _ops.stem_pool_forward(ps, xp, B, H, W, n_exp, out_pad)
    torch.cuda.synchronize()
    with torch.no_grad():
        # the reference pools bf16-rounded activations (autocast); max commutes with the rounding
        refs = [F.relu(bn(c(img))) for c, bn in zip(convs, bns)]
        ref_pool = torch.cat([F.max_pool2d(r, 3, 2, 1) for r in refs[:n_exp]], 0)
    y = pooled
    if out_pad:
        assert (y[:, 0] == 0).all() and (y[:, -1] == 0).all() and (y[:, :, 0] == 0).all() and (y[:, :, -1] == 0).all()
        y = y[:, 1:-1, 1:-1]
    y = y.float().permute(0, 3, 1, 2)
    assert y.shape == ref_pool.shape
    assert rel_err(y, ref_pool) < 8e-3 and rel_l2(y, ref_pool) < 4e-3, (rel_err(y, ref_pool), rel_l2(y, ref_pool))
    if with_policy:
        yp = rest[0].float().permute(0, 3, 1, 2)
        assert yp.shape == refs[-1].shape
        assert rel_err(yp, refs[-1]) < 8e-3, rel_err(yp, refs[-1])


def _mk_conv_bn(Cin, Cout, k, s, p, g, n=1, bias=False):
    convs, bns = [], []
    for _ in range(n):
        c = nn.Conv2d(Cin, Cout, k, s, p, bias=bias)
        b = nn.BatchNorm2d(Cout)
        with torch.no_grad():
            c.weight.copy_((torch.randn(c.weight.shape, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).bfloat16().float())
            b.weight.copy_(1 + 0.1 * torch.randn(Cout, generator=g))
            b.bias.copy_(0.1 * torch.randn(Cout, generator=g))
            b.running_mean.copy_(0.1 * torch.randn(Cout, generator=g))
            b.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
        convs.append(c.to(DEV))
        bns.append(b.to(DEV).eval())
    return convs, bns


def _pad_nhwc(x_nchw, dtype):
    """NCHW fp32 -> physically padded NHWC [N,H+2,W+2,C] with a zero border."""
    return F.pad(x_nchw.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).to(dtype).contiguous()


FLAT_SHAPES = [
    # G, B, H, W, C, N, residual
    (1, 2, 16, 16, 64, 64, False),     # resident weights, 3 tiles
    (1, 2, 16, 16, 64, 64, True),
    (3, 2, 64, 64, 64, 64, True),      # layer1 geometry, grouped experts
    (1, 3, 32, 32, 128, 128, True),    # layer2: 2 chunks, streamed weights
    (2, 1, 56, 56, 64, 64, False),     # 224-px geometry
    (1, 5, 14, 14, 128, 128, True),
    (1, 40, 16, 16, 64, 64, False),    # many tiles per CTA: both accumulators, ring wrap-around
    (1, 2, 8, 8, 64, 128, False),      # Cin != Cout
    (2, 3, 13, 11, 64, 64, True),      # odd sizes: ragged last tile, group boundary inside a kw-fused sub-tile
    (1, 1, 4, 4, 64, 64, False),       # a single partial tile
]


@pytest.mark.timeout(180)
@pytest.mark.parametrize("kw3", ["0", "1", "pair"])
@pytest.mark.parametrize("shape", FLAT_SHAPES)
def test_conv_flat_bf16(shape, kw3, monkeypatch):
    """Halo-reuse 3x3/s1 kernel on physically padded activations (csrc/conv_flat.cu).  kw3=1: the opt-in
    64->64 variant that multiplies two horizontal taps per MMA and shifts rows in the epilogue; pair: CTA pairs
    issuing M=256 tcgen05.mma.cta_group::2 with the weight tiles split over the two CTAs."""
    from automoe_b200 import _ops
    G, B, H, W, C, N, residual = shape
    if kw3 == "1" and (C != 64 or N != 64):
        pytest.skip("kw-fused kernel is 64 -> 64 channels only")
    monkeypatch.setenv("AMOE_FLAT_KW3", "1" if kw3 == "1" else "0")
    monkeypatch.setenv("AMOE_FLAT_PAIR", "1" if kw3 == "pair" else "0")
    g = torch.Generator().manual_seed(6)
    convs, bns = _mk_conv_bn(C, N, 3, 1, 1, g, n=G)
    x = torch.randn((G * B, C, H, W), generator=g).bfloat16().float().to(DEV)
    res = torch.randn((G * B, N, H, W), generator=g).bfloat16().float().to(DEV) if residual else None
    pc = _ops.pack_conv(convs, bns, torch.bfloat16, torch.device(DEV), relu=True)
    assert _ops.flat_supported(pc, H, W, torch.bfloat16)
    xp = _pad_nhwc(x, torch.bfloat16)
    rp = _pad_nhwc(res, torch.bfloat16) if residual else None
    _ops.walk_reset(DEV, 0)
    y = _ops.conv3x3_flat(pc, xp, B, H, W, residual=rp)
    y_rev = _ops.conv3x3_flat(pc, xp, B, H, W, residual=rp)      # the next launch walks its tiles back to front
    torch.cuda.synchronize()
    assert torch.equal(y, y_rev)
    with torch.no_grad():
        ref = torch.cat([F.relu(bns[i](convs[i](x[i * B:(i + 1) * B])) + (res[i * B:(i + 1) * B] if residual else 0))
                         for i in range(G)], 0)
    yi = y[:, 1:-1, 1:-1, :].float().permute(0, 3, 1, 2)
    assert (y[:, 0] == 0).all() and (y[:, -1] == 0).all() and (y[:, :, 0] == 0).all() and (y[:, :, -1] == 0).all()
    assert rel_err(yi, ref) < 8e-3 and rel_l2(yi, ref) < 4e-3, (rel_err(yi, ref), rel_l2(yi, ref))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", [(64, 128, 3, 2, 1, 32, 32), (64, 128, 1, 2, 0, 32, 32), (128, 256, 3, 2, 1, 16, 16),
                                  (64, 64, 3, 1, 1, 16, 16)])
def test_conv_tc_padded_layouts(case):
    """Per-tap tcgen05 kernel reading a physically padded input and/or writing a padded output."""
    from automoe_b200 import _ops
    Cin, Cout, k, s, p, H, W = case
    B = 3
    g = torch.Generator().manual_seed(8)
    convs, bns = _mk_conv_bn(Cin, Cout, k, s, p, g)
    x = torch.randn((B, Cin, H, W), generator=g).bfloat16().float().to(DEV)
    pc = _ops.pack_conv(convs, bns, torch.bfloat16, torch.device(DEV), relu=True)
    xp = _pad_nhwc(x, torch.bfloat16)
    y_pp = _ops.conv2d(pc, xp, B, H, W, in_pad=1, out_pad=1, zero_border=True)
    y_pu = _ops.conv2d(pc, xp, B, H, W, in_pad=1, out_pad=0)
    y_up = _ops.conv2d(pc, x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous(), B, H, W, in_pad=0, out_pad=1, zero_border=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = F.relu(bns[0](convs[0](x)))
    for y in (y_pp[:, 1:-1, 1:-1], y_pu, y_up[:, 1:-1, 1:-1]):
        yy = y.float().permute(0, 3, 1, 2)
        assert yy.shape == ref.shape
        assert rel_err(yy, ref) < 8e-3, rel_err(yy, ref)
    assert (y_pp[:, 0] == 0).all() and (y_pp[:, :, -1] == 0).all()


def test_maxpool():
    from automoe_b200 import _ops
    for dtype in (torch.float32, torch.bfloat16):
        for (N, H, W, C) in [(3, 16, 16, 64), (2, 9, 7, 24), (1, 14, 14, 3)]:
            x = torch.randn(N, C, H, W, device=DEV).to(dtype)
            ref = F.max_pool2d(x.float(), 3, 2, 1)
            y = _ops.maxpool3x3s2(x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2).float()
            assert torch.equal(y, ref)
            yp = _ops.maxpool3x3s2(x.permute(0, 2, 3, 1).contiguous(), out_pad=1)
            assert torch.equal(yp[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float(), ref)
            assert (yp[:, 0] == 0).all() and (yp[:, -1] == 0).all() and (yp[:, :, 0] == 0).all() and (yp[:, :, -1] == 0).all()


def test_image_to_nhwc():
    from automoe_b200 import _ops
    img = torch.randn(3, 3, 10, 12, device=DEV)
    for dtype, cp in [(torch.float32, 4), (torch.bfloat16, 4), (torch.bfloat16, 8)]:
        y = _ops.image_to_nhwc(img, cp, dtype)
        assert torch.equal(y[..., :3].float(), img.permute(0, 2, 3, 1).to(dtype).float())
        assert (y[..., 3:] == 0).all()


def test_head1x1_pool_and_upsample():
    from automoe_b200 import _ops
    g = torch.Generator().manual_seed(3)
    for dtype in (torch.float32, torch.bfloat16):
        B, h, w, Cin, N = 3, 8, 8, 256, 19
        x = torch.randn((B, h, w, Cin), generator=g).to(DEV).to(dtype)
        wt = (torch.randn((N, Cin), generator=g) / 16).to(DEV)
        b = torch.randn(N, generator=g).to(DEV)
        pooled = torch.zeros((B, N + 5), device=DEV)
        low = _ops.head1x1_pool(x, wt, b, pooled, 5)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt[:, :, None, None], b)        # [B,N,h,w]
        assert rel_err(low.permute(0, 3, 1, 2), ref) < 1e-5
        assert rel_err(pooled[:, 5:], ref.mean(dim=(2, 3))) < 1e-5
        assert (pooled[:, :5] == 0).all()
        for (H, W) in [(256, 256), (h, w), (100, 52)]:
            up = _ops.upsample_bilinear_nchw(low, H, W, dtype)
            upr = F.interpolate(low.permute(0, 3, 1, 2).contiguous(), size=(H, W), mode="bilinear", align_corners=False)
            tol = 1e-5 if dtype == torch.float32 else 4e-3
            assert rel_err(up.float(), upr) < tol, (H, W, rel_err(up.float(), upr))
        m = _ops.mean_hw_nchw(upr.to(dtype))
        assert rel_err(m, upr.to(dtype).float().mean(dim=(2, 3))) < 1e-5


def _small_model():
    from _util import build_b200_model
    return build_b200_model(DEV, "fp32")


def test_gate_kernel_vs_oracle():
    """Fused gate kernel on given pooled logits + vehicle state: every output against the oracle,
    top-1 routing bit-exact."""
    from automoe_b200 import _ops
    from oracle import automoe_oracle as O
    m, sd = _small_model()
    sd = {k: v.to(DEV) for k, v in sd.items()}
    g = torch.Generator().manual_seed(5)
    for B in (1, 3, 4, 67, 256):
        pooled = torch.randn((B, 36), generator=g).to(DEV) * 3
        state = torch.cat([torch.rand((B, 1), generator=g) * 30, torch.rand((B, 3), generator=g) - 0.5], 1).to(DEV)
        n_ch = [14, 19, 3]
        out = _ops.gate(state, pooled, m._gate_params(torch.device(DEV), n_ch), n_ch, 64, 128, 1.0)
        # oracle on the same pooled statistics
        ctx = O.context_extractor(state, sd)
        feats, off = [], 0
        for i, n in enumerate(n_ch):
            p = f"expert_extractors.extractors.{i}.feature_extractor"
            v = F.relu(F.linear(pooled[:, off:off + n], sd[p + ".2.weight"], sd[p + ".2.bias"]))
            v = F.linear(v, sd[p + ".5.weight"], sd[p + ".5.bias"])
            feats.append(F.layer_norm(v, (256,), sd[p + ".6.weight"], sd[p + ".6.bias"]))
            off += n
        ref = O.gating_network(feats, ctx, sd)
        assert rel_err(out["context"], ctx) < 1e-5
        for i in range(3):
            assert rel_err(out["features"][i], feats[i]) < 1e-5
            assert rel_err(out["processed"][i], ref["processed_expert_outputs"][i]) < 1e-5
        assert rel_err(out["gate_logits"], ref["gate_logits"]) < 1e-5
        assert rel_err(out["weights"], ref["expert_weights"]) < 1e-5
        assert rel_err(out["combined"], ref["combined_output"]) < 1e-5
        assert torch.equal(out["weights"].argmax(1), ref["expert_weights"].argmax(1))
        assert torch.allclose(out["weights"].sum(1), torch.ones(B, device=DEV), atol=1e-6)


def test_submodules_standalone_vs_oracle():
    """The reference's unit tests call sub-modules directly (tests/test_gating_network.py:25-156)."""
    from oracle import automoe_oracle as O
    m, sd = _small_model()
    sd = {k: v.to(DEV) for k, v in sd.items()}
    g = torch.Generator().manual_seed(9)
    B = 5
    ctx = torch.randn((B, 64), generator=g).to(DEV)
    feats = [torch.randn((B, 256), generator=g).to(DEV) for _ in range(3)]
    out = m.gating_network(feats, ctx)
    ref = O.gating_network(feats, ctx, sd)
    for k in ("combined_output", "expert_weights", "gate_logits"):
        assert rel_err(out[k], ref[k]) < 1e-5, k
    assert rel_err(m.gating_network.get_expert_weights(ctx), O.gating_network([None] * 3, ctx, sd, context_only=True)["expert_weights"]) < 1e-5
    seg = torch.randn((B, 19, 32, 32), generator=g).to(DEV)
    f = m.expert_extractors.extractors[1](seg)
    assert rel_err(f, O.extractor(seg, sd, "expert_extractors.extractors.1", {"type": "segmentation"})) < 1e-5
    det = {"class_logits": torch.randn((B, 10, 8, 8), generator=g).to(DEV), "bbox_deltas": torch.randn((B, 4, 8, 8), generator=g).to(DEV)}
    f = m.expert_extractors.extractors[0](det)
    assert rel_err(f, O.extractor(det, sd, "expert_extractors.extractors.0", {"type": "detection"})) < 1e-5
    st = torch.rand((B, 4), generator=g).to(DEV)
    c = m.context_extractor(st[:, 0:1], st[:, 1:2], st[:, 2:3], st[:, 3:4])
    assert rel_err(c, O.context_extractor(st, sd)) < 1e-5


def test_policy_head_vs_oracle():
    from oracle import automoe_oracle as O
    m, sd = _small_model()
    sd = {k: v.to(DEV) for k, v in sd.items()}
    g = torch.Generator().manual_seed(2)
    for B, H in ((1, 64), (5, 96)):
        img = torch.randn((B, 3, H, H), generator=g).to(DEV)
        ctx = torch.randn((B, 256), generator=g).to(DEV)
        out = m.policy_head(img, context=ctx)
        ref = O.policy_head(img, ctx, sd)
        assert rel_err(out["waypoints"], ref["waypoints"]) < 1e-4
        assert rel_err(out["speed"], ref["speed"]) < 1e-4
        m.policy_head.precision = "bf16"
        outb = m.policy_head(img, context=ctx)
        m.policy_head.precision = "auto"
        assert rel_err(outb["waypoints"], ref["waypoints"]) < 2e-2


@pytest.mark.parametrize("B", [64, 100, 256])
def test_mlp_cluster_kernels_match_single_cta(B, monkeypatch):
    """Large batches run the gate / policy-head MLPs on clusters of 8 CTAs (output rows split across the
    cluster, activations all-gathered through distributed shared memory).  Same fp32 arithmetic with a
    different summation order: equal to the single-CTA kernels to rounding, identical top-1 routing
    (ragged last cluster included)."""
    from automoe_b200 import _ops
    m, _ = _small_model()
    g = torch.Generator().manual_seed(17)
    pooled = (torch.randn((B, 36), generator=g) * 3).to(DEV)
    state = torch.cat([torch.rand((B, 1), generator=g) * 30, torch.rand((B, 3), generator=g) - 0.5], 1).to(DEV)
    n_ch = [14, 19, 3]
    prm = m._gate_params(torch.device(DEV), n_ch)
    x = torch.randn((B, 4, 4, 256), generator=g).to(DEV)
    cvec = torch.randn((B, 256), generator=g).to(DEV)
    pp = m.policy_head._pack(torch.float32, torch.device(DEV))["flat"]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("AMOE_MLP_CLUSTER", mode)
        gate = _ops.gate(state, pooled, prm, n_ch, 64, 128, 1.0)
        wp32, sp32 = _ops.policy_head(x, cvec, pp, 512, 256, 512, 10)
        wp16, sp16 = _ops.policy_head(x.bfloat16(), cvec, pp, 512, 256, 512, 10)
        torch.cuda.synchronize()
        res[mode] = (gate, wp32, sp32, wp16, sp16)
    for k in ("context", "features", "processed", "gate_logits", "weights", "combined"):
        assert rel_err(res["1"][0][k], res["0"][0][k]) < 5e-6, (k, rel_err(res["1"][0][k], res["0"][0][k]))
    assert torch.equal(res["1"][0]["weights"].argmax(1), res["0"][0]["weights"].argmax(1))
    for i in range(1, 5):
        assert rel_err(res["1"][i], res["0"][i]) < 5e-6, i


@pytest.mark.parametrize("B", [16, 37, 256])
def test_mlp_tf32_kernels_vs_fp32_kernels(B):
    """bf16 inference mode runs the gate / policy-head layers 16 frames per CTA on mma.sync TF32 (bf16 weights,
    fp32 activations truncated to TF32, fp32 accumulate).  Against the fp32 CUDA-core kernels fed the SAME
    bf16-rounded weights: every output within 2e-3 relative (activation truncation only); against the fp32
    weights: within 1e-2 (the bf16 tolerance - this is the weight rounding the reference's autocast applies too);
    softmax weights sum to one, identical top-1 routing wherever the fp32 logit gap exceeds that noise; ragged
    last CTA included.  The grid-wide NHWC mean kernel that feeds the head is checked against torch."""
    from automoe_b200 import _ops
    m, _ = _small_model()
    g = torch.Generator().manual_seed(23)
    pooled = (torch.randn((B, 36), generator=g) * 3).to(DEV)
    state = torch.cat([torch.rand((B, 1), generator=g) * 30, torch.rand((B, 3), generator=g) - 0.5], 1).to(DEV)
    n_ch = [14, 19, 3]
    prm, prm16 = m._gate_params(torch.device(DEV), n_ch, bf16_copy=True)
    ref = _ops.gate(state, pooled, prm, n_ch, 64, 128, 1.0)
    ref16 = _ops.gate(state, pooled, prm16.float(), n_ch, 64, 128, 1.0)     # fp32 kernel, bf16-rounded parameters
    out = _ops.gate(state, pooled, prm, n_ch, 64, 128, 1.0, params_bf16=prm16)
    for k in ("context", "features", "processed", "gate_logits", "weights", "combined"):
        assert rel_err(out[k], ref[k]) < 1e-2, (k, rel_err(out[k], ref[k]))
    # weights of the MMA layers are the bf16 copy, biases / LayerNorm / tiny layers the fp32 buffer: compare the
    # end of the chain with both references
    assert min(rel_err(out["combined"], ref16["combined"]), rel_err(out["combined"], ref["combined"])) < 5e-3
    assert torch.allclose(out["weights"].sum(1), torch.ones(B, device=DEV), atol=1e-6)
    top2 = ref["gate_logits"].topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2e-2 * ref["gate_logits"].abs().max()
    assert safe.float().mean() > 0.5
    assert torch.equal(out["weights"].argmax(1)[safe], ref["weights"].argmax(1)[safe])

    x = torch.randn((B, 4, 4, 256), generator=g).to(DEV).bfloat16()
    cvec = torch.randn((B, 256), generator=g).to(DEV)
    pk = m.policy_head._pack(torch.float32, torch.device(DEV))
    pp, pp16 = pk["flat"], pk["flat16"]
    pooled_x = _ops.mean_hw_nhwc(x)
    assert rel_err(pooled_x, x.float().mean(dim=(1, 2))) < 1e-6
    wp_ref, sp_ref = _ops.policy_head(x, cvec, pp, 512, 256, 512, 10)
    wp, sp = _ops.policy_head(x, cvec, pp, 512, 256, 512, 10, params_bf16=pp16)
    assert rel_err(wp, wp_ref) < 1e-2, rel_err(wp, wp_ref)
    assert rel_err(sp, sp_ref) < 1e-2, rel_err(sp, sp_ref)
    # same bf16-rounded weights through the fp32 kernel: only the TF32 truncation of the activations differs
    # (biases differ by their bf16 rounding there, hence not tighter than 3e-3)
    wp_r16, sp_r16 = _ops.policy_head(x, cvec, pp16.float(), 512, 256, 512, 10)
    assert rel_err(wp, wp_r16) < 3e-3, rel_err(wp, wp_r16)
    assert rel_err(sp, sp_r16) < 3e-3, rel_err(sp, sp_r16)
    # fp32 feature map through the tensor-core head (pooling inside the kernel)
    wp32, sp32 = _ops.policy_head(x.float(), cvec, pp, 512, 256, 512, 10, params_bf16=pp16)
    assert rel_err(wp32, wp) < 1e-3 and rel_err(sp32, sp) < 1e-3
