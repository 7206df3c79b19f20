"""Drop-in for the model-facing functions of inference/run_automoe.py (build_image_transform :25-31, model_infer
:34-53, load_model :144-156) with the per-frame input transform moved onto the GPU.

The reference resizes and normalises every camera frame on the CPU (PIL + torchvision) and uploads 12 bytes per
pixel; here the uint8 HWC frame is uploaded as it is (3 bytes per pixel) and `amoe_resample_u8_fwd` (Pillow's 8-bit
bilinear resize, bit-exact) + `amoe_stage_u8_hwc_fwd` (ToTensor + Normalize, bit-exact in fp32, written straight into
the bf16 NHWC layout of the tensor-core stem) run on the device.  The CARLA client, PID controller and the driving
loop of the reference script are control plane and out of scope; they call these three functions unchanged.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _ops
from ..models.automoe import create_automoe_model


class DeviceImageTransform:
    """What build_image_transform returns: callable like the reference's T.Compose on one uint8 HWC frame (result
    [3,h,w] fp32 on the frame's device, bit-identical to the CPU transform), and `.stage(frames)` for the batched
    device path used by model_infer (uint8 [B,H,W,3] -> resized uint8 [B,h,w,3], normalisation left to the model's
    fused staging kernel)."""

    def __init__(self, target_hw: Tuple[int, int] = (256, 256), device=None):
        self.target_hw = (int(target_hw[0]), int(target_hw[1]))
        self.device = device
        self.mean, self.std = _ops.IMAGENET_MEAN, _ops.IMAGENET_STD

    def _to_device(self, image_rgb, device) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(image_rgb)) if isinstance(image_rgb, np.ndarray) else image_rgb
        if t.dtype != torch.uint8:
            raise TypeError(f"camera frames must be uint8 RGB, got {t.dtype}")
        if t.dim() == 3:
            t = t.unsqueeze(0)
        dev = torch.device(device if device is not None else (self.device or "cuda"))
        return t.to(dev, non_blocking=True)

    def stage(self, image_rgb, device=None) -> torch.Tensor:
        frames = self._to_device(image_rgb, device)
        return _ops.resize_u8_bilinear(frames, *self.target_hw)

    def __call__(self, image_rgb, device=None) -> torch.Tensor:
        single = image_rgb.ndim == 3
        out = _ops.normalize_u8_nchw(self.stage(image_rgb, device), self.mean, self.std)
        return out[0] if single else out


def build_image_transform(target_hw: Tuple[int, int] = (256, 256)) -> DeviceImageTransform:
    return DeviceImageTransform(target_hw)


@torch.no_grad()
def model_infer(model: nn.Module, image_rgb: np.ndarray, last_speed_kmh: float, device: torch.device,
                img_tf=None) -> Dict[str, torch.Tensor]:
    """image_rgb: [H,W,3] uint8 (or a batch [B,H,W,3]); same batch dict and autocast as run_automoe.py:41-52."""
    device = torch.device(device)
    tf = img_tf if isinstance(img_tf, DeviceImageTransform) else build_image_transform((256, 256))
    frames = tf.stage(image_rgb, device)                     # uint8 [B,h,w,3] on the device
    B = frames.shape[0]
    speed = torch.full((B, 1), float(last_speed_kmh), dtype=torch.float32, device=device)
    batch: Dict[str, Any] = {
        'image': frames,
        'speed': speed,
        'steering': torch.zeros(B, 1, device=device),
        'throttle': torch.zeros(B, 1, device=device),
        'brake': torch.zeros(B, 1, device=device),
    }
    with torch.autocast(device_type='cuda', enabled=True):
        pred = model(batch)
    return pred


def load_model(model_config_path: str, checkpoint_path: str, device: torch.device) -> nn.Module:
    cfg = json.loads(Path(model_config_path).read_text())
    model = create_automoe_model(cfg, device)
    state = torch.load(checkpoint_path, map_location=device)
    state_dict = state.get('model_state_dict', state)
    if any(k.startswith('module.') for k in state_dict.keys()):      # DDP prefixes
        state_dict = {k[len('module.'):]: v for k, v in state_dict.items()}
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    if missing or unexpected:
        print(f"Loaded with relaxed matching. Missing={len(missing)} Unexpected={len(unexpected)}")
    model.eval()
    return model
