#!/usr/bin/env python
"""Times the fused stem + max-pool kernel (csrc/stem_tc.cu) alone at the bench geometry (3 expert stems + policy conv1,
256 frames of 256x256) under several environment settings, CUDA events, 20 launches.

    python tools/stem_bench.py "AMOE_STEM_DBG=0" "AMOE_STEM_DBG=1" ...
"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automoe_b200 import _ops  # noqa: E402

DEV = torch.device("cuda:0")


def main(settings, B=256, H=256, W=256):
    torch.manual_seed(0)
    convs = [nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(DEV) for _ in range(3)] + [nn.Conv2d(3, 32, 5, 2, 2, bias=True).to(DEV)]
    bns = [nn.BatchNorm2d(c.weight.shape[0]).to(DEV).eval() for c in convs]
    ps = _ops.pack_stem(convs, bns, DEV, relu=True)
    imgs = [torch.randn((B, 3, H, W), device=DEV) for _ in range(2)]
    xps = [_ops.stage_image_stem(i) for i in imgs]
    pooled, rest = _ops.stem_pool_forward(ps, xps[0], B, H, W, 3, 1)
    torch.cuda.synchronize()
    ref = pooled.clone()
    flop = 2.0 * ps.true_macs_per_px * B * (H // 2) * (W // 2)
    for st in settings:
        for kv in st.split():
            k, v = kv.split("=")
            os.environ[k] = v
        for i in range(3):
            _ops.stem_pool_forward(ps, xps[i & 1], B, H, W, 3, 1, pooled=pooled, rest=rest)
        _ops.stem_pool_forward(ps, xps[0], B, H, W, 3, 1, pooled=pooled, rest=rest)
        torch.cuda.synchronize()
        same = bool((pooled == ref).all())
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            _ops.stem_pool_forward(ps, xps[i & 1], B, H, W, 3, 1, pooled=pooled, rest=rest)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print("%-40s %7.1f us  %6.1f useful TFLOP/s  bit-identical=%s" % (st, us, flop / us / 1e6, same), flush=True)
        for kv in st.split():
            os.environ.pop(kv.split("=")[0], None)


if __name__ == "__main__":
    main(sys.argv[1:] or ["AMOE_STEM_DBG=0"])
