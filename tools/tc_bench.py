#!/usr/bin/env python
"""Times the per-tap tcgen05 convolution (csrc/conv_tc.cu) alone on the bench geometry's layer3 / layer4 / head / policy
shapes (3 experts x 256 frames) under several environment settings, CUDA events, 20 launches.

    python tools/tc_bench.py "AMOE_TC_TAIL_SPLIT=0" "AMOE_TC_TAIL_SPLIT=1"
"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automoe_b200 import _ops  # noqa: E402

DEV = torch.device("cuda:0")
CASES = [  # name, G, B, H, W, Cin, Cout, k, stride, residual
    ("layer3 3x3", 3, 256, 16, 16, 256, 256, 3, 1, False),
    ("layer3 3x3+res", 3, 256, 16, 16, 256, 256, 3, 1, True),
    ("layer4 3x3", 3, 256, 8, 8, 512, 512, 3, 1, False),
    ("layer4 3x3+res", 3, 256, 8, 8, 512, 512, 3, 1, True),
    ("head 3x3", 3, 256, 8, 8, 512, 256, 3, 1, False),
    ("layer4 entry s2", 3, 256, 16, 16, 256, 512, 3, 2, False),
    ("layer2 entry s2", 3, 256, 64, 64, 64, 128, 3, 2, False),
    ("policy conv3 s2", 1, 256, 64, 64, 64, 128, 3, 2, False),
    ("policy conv4 s2", 1, 256, 32, 32, 128, 256, 3, 2, False),
]


def main(settings):
    torch.manual_seed(0)
    only = os.environ.get("TC_BENCH_ONLY")
    for name, G, B, H, W, C, N, k, s, res in CASES:
        if only and only not in name:
            continue
        convs = [nn.Conv2d(C, N, k, s, k // 2, bias=False).to(DEV) for _ in range(G)]
        bns = [nn.BatchNorm2d(N).to(DEV).eval() for _ in range(G)]
        pc = _ops.pack_conv(convs, bns, torch.bfloat16, DEV, relu=True)
        xs = [torch.randn((G * B, H, W, C), device=DEV).bfloat16() for _ in range(2)]
        Ho, Wo = H // s, W // s
        rs = [torch.randn((G * B, Ho, Wo, N), device=DEV).bfloat16() for _ in range(2)] if res else [None, None]
        flop = 2.0 * k * k * C * N * G * B * Ho * Wo
        ref = None
        for st in settings:
            for kv in st.split():
                a, v = kv.split("=")
                os.environ[a] = v
            for i in range(3):
                y = _ops.conv2d(pc, xs[i & 1], B, H, W, residual=rs[i & 1])
            y = _ops.conv2d(pc, xs[0], B, H, W, residual=rs[0])
            torch.cuda.synchronize()
            if ref is None:
                ref = y.clone()
            same = bool((y == ref).all())
            n = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                _ops.conv2d(pc, xs[i & 1], B, H, W, residual=rs[i & 1])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            print("%-16s %-36s %7.1f us  %6.1f TFLOP/s  bit-identical-to-first=%s" % (name, st, us, flop / us / 1e6, same), flush=True)
            for kv in st.split():
                os.environ.pop(kv.split("=")[0], None)


if __name__ == "__main__":
    main(sys.argv[1:] or ["AMOE_NONE=0"])
