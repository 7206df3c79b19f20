"""CPU: the staging oracle (oracle/stage_oracle.py) against the reference's golden outputs
(tests/golden/stage_*.npz, produced by the reference's build_image_transform = real Pillow + torchvision), against
Pillow itself where it is importable, and the product's host-side coefficient tables against the oracle's."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

from oracle import stage_oracle as SO
from oracle.synth import synth_u8_frame

GOLD = sorted((Path(__file__).parent / "golden").glob("stage_*.npz"))


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_oracle_reproduces_reference_transform_bit_exactly(path):
    g = np.load(path)
    (ih, iw), (oh, ow) = g["in_hw"], g["out_hw"]
    frame = synth_u8_frame(int(ih), int(iw), int(g["seed"]))
    out = SO.transform(frame, (int(oh), int(ow)))
    assert out.dtype == np.float32 and out.shape == (3, oh, ow)
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["sha256"])
    if "out" in g:
        assert np.array_equal(out, g["out"])
    else:
        assert np.array_equal(out[:, 100:116, 120:136], g["crop"]) and np.array_equal(out[:, 255, :], g["row"])


@pytest.mark.parametrize("ih,iw,oh,ow", [(64, 64, 64, 64), (50, 70, 20, 30), (33, 47, 64, 64), (600, 800, 256, 256), (17, 301, 5, 7)])
def test_oracle_resize_equals_pillow(ih, iw, oh, ow):
    PIL = pytest.importorskip("PIL.Image")
    frame = synth_u8_frame(ih, iw, 3)
    ref = np.asarray(PIL.fromarray(frame).resize((ow, oh), PIL.BILINEAR))
    assert np.array_equal(SO.resize_bilinear_u8(frame, (oh, ow)), ref)


def test_batched_oracle_equals_per_frame():
    frames = np.stack([synth_u8_frame(40, 56, s) for s in (1, 2, 3)])
    out = SO.transform(frames, (16, 24))
    for i in range(3):
        assert np.array_equal(out[i], SO.transform(frames[i], (16, 24)))


@pytest.mark.parametrize("n_in,n_out", [(800, 256), (600, 256), (256, 256), (20, 32), (53, 24), (7, 7), (301, 7)])
def test_product_coefficient_tables_equal_oracle(n_in, n_out):
    """Host logic of the product path (_ops.pil_bilinear_coeffs feeds amoe_resample_u8_fwd)."""
    from automoe_b200 import _ops
    bounds, coeffs, ksize = _ops.pil_bilinear_coeffs(n_in, n_out)
    xmin, cnt, kk = SO.coeffs_8bpc(n_in, n_out)
    assert ksize == kk.shape[1]
    assert np.array_equal(bounds[:, 0].numpy(), xmin) and np.array_equal(bounds[:, 1].numpy(), cnt)
    assert np.array_equal(coeffs.numpy(), kk)
