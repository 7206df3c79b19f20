"""GPU (-m gpu): input staging from camera bytes through the C-ABI against the oracle
(oracle/stage_oracle.py, pinned to the reference transform) and the reference's golden outputs: integer resize
bit-exact, normalisation bit-exact in fp32, staged bf16 frame = round-to-nearest of the fp32 value; and the model
fed uint8 frames gives exactly what it gives for the CPU-transformed fp32 tensor."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from _util import build_b200_model
from oracle import stage_oracle as SO
from oracle.synth import synth_u8_frame

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = sorted((Path(__file__).parent / "golden").glob("stage_*.npz"))


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_device_transform_reproduces_reference_golden(path):
    from automoe_b200.inference import build_image_transform
    g = np.load(path)
    (ih, iw), (oh, ow) = g["in_hw"], g["out_hw"]
    frame = synth_u8_frame(int(ih), int(iw), int(g["seed"]))
    out = build_image_transform((int(oh), int(ow)))(frame, DEV).cpu().numpy()
    assert out.shape == (3, oh, ow) and out.dtype == np.float32
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["sha256"])


@pytest.mark.parametrize("B,ih,iw,oh,ow", [(3, 50, 70, 20, 30), (2, 33, 47, 64, 64), (2, 600, 800, 256, 256), (5, 17, 301, 5, 7),
                                           (1, 64, 48, 64, 24), (0, 8, 8, 4, 4)])
def test_resize_u8_bit_exact(B, ih, iw, oh, ow):
    from automoe_b200 import _ops
    frames = np.stack([synth_u8_frame(ih, iw, 20 + i) for i in range(B)]) if B else np.zeros((0, ih, iw, 3), np.uint8)
    out = _ops.resize_u8_bilinear(torch.from_numpy(frames).to(DEV), oh, ow).cpu().numpy()
    assert out.shape == (B, oh, ow, 3)
    if B:
        assert np.array_equal(out, SO.resize_bilinear_u8(frames, (oh, ow)))


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (3, 64, 64), (2, 30, 46), (1, 8, 250)])
def test_stage_u8_stem_layout_and_values(B, H, W):
    """[B,H+6,Wpad,4] bf16: zero border (3 rows / 4 px), channel 3 == 1 everywhere, interior == bf16(reference fp32)."""
    from automoe_b200 import _ops
    rng = np.random.RandomState(5)
    frames = rng.randint(0, 256, size=(B, H, W, 3)).astype(np.uint8)
    frames[0, 0, :, :] = 0
    frames[0, 1, :, :] = 255
    ref = torch.from_numpy(SO.to_tensor_normalize(frames))                 # [B,3,H,W] fp32
    x = _ops.stage_u8_stem(torch.from_numpy(frames).to(DEV)).cpu()
    Wpad = _ops.stem_wpad(W)
    assert x.shape == (B, H + 6, Wpad, 4) and x.dtype == torch.bfloat16
    assert torch.equal(x[..., 3], torch.ones_like(x[..., 3]))
    inner = x[:, 3:3 + H, 4:4 + W, :3].permute(0, 3, 1, 2)
    assert torch.equal(inner, ref.to(torch.bfloat16))
    mask = torch.ones((H + 6, Wpad), dtype=torch.bool)
    mask[3:3 + H, 4:4 + W] = False
    assert float(x[:, mask][..., :3].float().abs().max()) == 0.0
    # same bits as staging the fp32 tensor the reference would have uploaded
    x2 = _ops.stage_image_stem(ref.to(DEV)).cpu()
    assert torch.equal(x, x2)
    # fp32 NCHW variant: bit-identical to ToTensor + Normalize
    assert torch.equal(_ops.normalize_u8_nchw(torch.from_numpy(frames).to(DEV)).cpu(), ref)


@pytest.mark.parametrize("autocast", [True, False])
def test_model_on_uint8_frames_equals_model_on_cpu_transformed_tensor(autocast):
    m, _ = build_b200_model(DEV, "auto")
    frames = np.stack([synth_u8_frame(256, 256, 40 + i) for i in range(4)])
    image = torch.from_numpy(SO.to_tensor_normalize(frames)).to(DEV)
    speed = torch.tensor([[3.0], [12.5], [0.0], [29.0]], device=DEV)
    z = torch.zeros(4, 1, device=DEV)
    ref_batch = dict(image=image, speed=speed, steering=z, throttle=z, brake=z)
    u8_batch = dict(ref_batch, image=torch.from_numpy(frames).to(DEV))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        a, b = m(ref_batch), m(u8_batch)
    for k in ("waypoints", "speed", "speed_seq", "expert_weights", "gate_logits", "combined_features"):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(a["expert_outputs"][1], b["expert_outputs"][1])
    assert b["expert_outputs"][1].shape == (4, 19, 256, 256)


def test_model_infer_mirror_with_camera_sized_frame():
    """model_infer(model, uint8 HWC camera frame, speed, device, img_tf) - the reference call site's signature - equals the
    model run on the oracle's CPU transform of the same frame."""
    from automoe_b200.inference import build_image_transform, model_infer
    m, _ = build_b200_model(DEV, "auto")
    frame = synth_u8_frame(600, 800, 77)
    pred = model_infer(m, frame, 21.0, torch.device(DEV), build_image_transform((256, 256)))
    image = torch.from_numpy(SO.transform(frame, (256, 256)))[None].to(DEV)
    z = torch.zeros(1, 1, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref = m(dict(image=image, speed=torch.tensor([[21.0]], device=DEV), steering=z, throttle=z, brake=z))
    for k in ("waypoints", "speed", "expert_weights"):
        assert torch.equal(pred[k], ref[k]), k
