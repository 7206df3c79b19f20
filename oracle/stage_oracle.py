"""ORACLE (test infrastructure only — never imported by the product path): CPU restatement of the
reference's per-frame input transform, inference/run_automoe.py:25-31

    T.Compose([T.ToPILImage(), T.Resize(target_hw, BILINEAR), T.ToTensor(), T.Normalize(mean, std)])

The resize arithmetic is NOT under /root/reference: it lives in Pillow (unpinned in the reference's
requirements.txt; 12.2.0 installed here), src/libImaging/Resample.c — `precompute_coeffs`,
`normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`,
`ImagingResampleInner` (horizontal pass first, 8-bit result, then the vertical pass).  Restated below in
numpy from the published algorithm; pinned by tests/golden/stage_*.npz, which
tests/golden/make_golden_stage.py produced by running the reference's own build_image_transform (i.e. the real
Pillow + torchvision) on seeded frames, and additionally against Pillow itself wherever it is importable
(tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c
MEAN = (0.485, 0.456, 0.406)         # run_automoe.py:30
STD = (0.229, 0.224, 0.225)


def _bilinear(x: float) -> float:
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def coeffs_8bpc(in_size: int, out_size: int):
    """precompute_coeffs (box = whole axis) + normalize_coeffs_8bpc.  -> (xmin[out], count[out], kk[out][ksize])."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale                      # bilinear support = 1
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int64)
    cnt = np.zeros(out_size, np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = 0 if lo < 0 else lo
        hi = int(center + support + 0.5)
        hi = in_size if hi > in_size else hi
        n = hi - lo
        w = [_bilinear((x + lo - center + 0.5) * inv) for x in range(n)]
        tot = 0.0
        for v in w:
            tot += v
        for x in range(n):
            v = w[x] / tot if tot != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass(a: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    """One 8-bit resample pass along `axis` of a uint8 array (ss0 = 1 << (PRECISION_BITS-1); clip8(ss0 >> bits))."""
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    xmin, cnt, kk = coeffs_8bpc(a.shape[0], out_size)
    out = np.empty((out_size,) + a.shape[1:], np.uint8)
    for o in range(out_size):
        acc = np.full(a.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(int(cnt[o])):
            acc += a[xmin[o] + t] * kk[o, t]
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, out_hw) -> np.ndarray:
    """img [..., H, W, 3] uint8 -> [..., out_h, out_w, 3] uint8, as PIL.Image.resize((w, h), BILINEAR)."""
    oh, ow = out_hw
    h_ax, w_ax = img.ndim - 3, img.ndim - 2
    if img.shape[w_ax] != ow:
        img = _pass(img, w_ax, ow)
    if img.shape[h_ax] != oh:
        img = _pass(img, h_ax, oh)
    return img


def to_tensor_normalize(img: np.ndarray, mean=MEAN, std=STD) -> np.ndarray:
    """[..., H, W, 3] uint8 -> [..., 3, H, W] float32: ToTensor (u8 -> f32, / 255) then Normalize ((x - mean) / std),
    every step one fp32 operation as torch executes it."""
    x = img.astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    x = (x - m) / s
    return np.moveaxis(x, -1, -3).astype(np.float32)


def transform(img: np.ndarray, target_hw=(256, 256)) -> np.ndarray:
    """build_image_transform(target_hw)(image_rgb), batched."""
    return to_tensor_normalize(resize_bilinear_u8(img, target_hw))
