#!/usr/bin/env python
"""Times the halo-reuse 3x3 convolutions (csrc/conv_flat.cu) alone at the bench geometry (3 experts x 256 frames) under
several environment settings, with CUDA events, inputs larger than L2 and rotated between launches.

    python tools/flat_bench.py "AMOE_FLAT_PAIR=0" "AMOE_FLAT_PAIR=1" "AMOE_FLAT_PAIR=1 AMOE_FLAT_PAIR_ISS=1"
"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automoe_b200 import _ops  # noqa: E402

DEV = torch.device("cuda:0")
CASES = [  # name, G, B, H, W, C, N, residual
    ("layer1", 3, 256, 64, 64, 64, 64, False),
    ("layer1+res", 3, 256, 64, 64, 64, 64, True),
    ("layer2", 3, 256, 32, 32, 128, 128, False),
    ("layer2+res", 3, 256, 32, 32, 128, 128, True),
]


def mk(G, C, N):
    convs = [nn.Conv2d(C, N, 3, 1, 1, bias=False).to(DEV) for _ in range(G)]
    bns = [nn.BatchNorm2d(N).to(DEV).eval() for _ in range(G)]
    return convs, bns


def main(settings):
    torch.manual_seed(0)
    rows = []
    for name, G, B, H, W, C, N, res in CASES:
        convs, bns = mk(G, C, N)
        pc = _ops.pack_conv(convs, bns, torch.bfloat16, DEV, relu=True)
        xs = [torch.randn((G * B, H + 2, W + 2, C), device=DEV).bfloat16() for _ in range(2)]
        for x in xs:
            x[:, 0] = 0; x[:, -1] = 0; x[:, :, 0] = 0; x[:, :, -1] = 0
        rs = [torch.randn((G * B, H + 2, W + 2, N), device=DEV).bfloat16() for _ in range(2)] if res else [None, None]
        out = torch.empty((G * B, H + 2, W + 2, N), device=DEV, dtype=torch.bfloat16)
        flop = 2.0 * 9 * C * N * G * B * H * W
        ref = None
        for st in settings:
            for kv in st.split():
                k, v = kv.split("=")
                os.environ[k] = v
            for i in range(3):
                _ops.conv3x3_flat(pc, xs[i & 1], B, H, W, residual=rs[i & 1], out=out)
            torch.cuda.synchronize()
            y = _ops.conv3x3_flat(pc, xs[0], B, H, W, residual=rs[0]).clone()
            if ref is None:
                ref = y
            same = bool((y == ref).all())
            n = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                _ops.conv3x3_flat(pc, xs[i & 1], B, H, W, residual=rs[i & 1], out=out)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            rows.append("%-12s %-44s %7.1f us  %6.1f TFLOP/s  bit-identical-to-first=%s" % (name, st, us, flop / us / 1e6, same))
            print(rows[-1], flush=True)
            for kv in st.split():
                os.environ.pop(kv.split("=")[0], None)
    return rows


if __name__ == "__main__":
    main(sys.argv[1:] or ["AMOE_FLAT_PAIR=0", "AMOE_FLAT_PAIR=1"])
