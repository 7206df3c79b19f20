#!/usr/bin/env python
"""Times the ResNet stage-entry dual launch (3x3/s2 conv1 + 1x1/s2 downsample, csrc/conv_tc.cu) alone at the bench geometry
(3 experts x 256 frames, physically padded input and output) under several environment settings.

    python tools/entry_bench.py "AMOE_NONE=0" "AMOE_TC_DBG=1"
"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automoe_b200 import _ops  # noqa: E402

DEV = torch.device("cuda:0")
CASES = [("layer2 entry", 64, 128, 64), ("layer3 entry", 128, 256, 32), ("layer4 entry", 256, 512, 16)]


def main(settings, G=3, B=256):
    torch.manual_seed(0)
    for name, C, N, H in CASES:
        convs = [nn.Conv2d(C, N, 3, 2, 1, bias=False).to(DEV) for _ in range(G)]
        bns = [nn.BatchNorm2d(N).to(DEV).eval() for _ in range(G)]
        ds = [nn.Conv2d(C, N, 1, 2, 0, bias=False).to(DEV) for _ in range(G)]
        dbn = [nn.BatchNorm2d(N).to(DEV).eval() for _ in range(G)]
        pc = _ops.pack_conv(convs, bns, torch.bfloat16, DEV, relu=True)
        pd = _ops.pack_conv(ds, dbn, torch.bfloat16, DEV, relu=False)
        in_pad = 1 if C <= 128 else 0     # layer1/2 activations are stored with a physical border
        out_pad = 1 if N <= 128 else 0
        xs = [torch.randn((G * B, H + 2 * in_pad, H + 2 * in_pad, C), device=DEV).bfloat16() for _ in range(2)]
        flop = 2.0 * 10 * C * N * G * B * (H // 2) * (H // 2)
        for st in settings:
            for kv in st.split():
                a, v = kv.split("=")
                os.environ[a] = v
            for i in range(3):
                _ops.conv2d_dual(pc, pd, xs[i & 1], B, H, H, in_pad=in_pad, out_pad=out_pad)
            torch.cuda.synchronize()
            n = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                _ops.conv2d_dual(pc, pd, xs[i & 1], B, H, H, in_pad=in_pad, out_pad=out_pad)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            print("%-14s %-36s %7.1f us  %6.1f TFLOP/s" % (name, st, us, flop / us / 1e6), flush=True)
            for kv in st.split():
                os.environ.pop(kv.split("=")[0], None)


if __name__ == "__main__":
    main(sys.argv[1:] or ["AMOE_NONE=0"])
