"""Golden vectors for the input-staging row: the UNMODIFIED reference transform
`inference.run_automoe.build_image_transform` (real Pillow resize + torchvision ToTensor/Normalize) applied to
seeded uint8 frames.

    python tests/golden/make_golden_stage.py

Needs /root/reference (build container only).  Inputs are regenerated at test time from the seed
(`oracle.synth.synth_u8_frame`); outputs are stored in full for the small cases and as crop + SHA-256 for the camera-sized one.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import PIL
import torch
import torchvision

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from inference.run_automoe import build_image_transform  # noqa: E402  (reference)

from oracle.synth import synth_u8_frame as stage_frame  # noqa: E402

OUT = Path(__file__).resolve().parent
VERS = np.array([f"torch {torch.__version__}", f"torchvision {torchvision.__version__}", f"Pillow {PIL.__version__}"])

# (name, in_h, in_w, out_h, out_w, seed)
CASES = [
    ("stage_identity_48", 48, 48, 48, 48, 11),
    ("stage_down_37x53_to_16x24", 37, 53, 16, 24, 12),
    ("stage_down_60x80_to_32x32", 60, 80, 32, 32, 13),
    ("stage_up_20x28_to_32x40", 20, 28, 32, 40, 14),
    ("stage_mixed_40x24_to_24x40", 40, 24, 24, 40, 15),
    ("stage_camera_600x800_to_256x256", 600, 800, 256, 256, 16),
]


def main():
    for name, ih, iw, oh, ow, seed in CASES:
        frame = stage_frame(ih, iw, seed)
        tf = build_image_transform((oh, ow))
        out = tf(frame).numpy()                      # [3, oh, ow] fp32
        assert out.shape == (3, oh, ow) and out.dtype == np.float32
        d = dict(in_hw=np.array([ih, iw]), out_hw=np.array([oh, ow]), seed=np.array(seed), versions=VERS,
                 sha256=np.array(hashlib.sha256(out.tobytes()).hexdigest()))
        if oh * ow <= 64 * 64:
            d["out"] = out
        else:
            d["crop"] = out[:, 100:116, 120:136].copy()
            d["row"] = out[:, 255, :].copy()
        np.savez_compressed(OUT / f"{name}.npz", **d)
        print(name, out.shape, d["sha256"])


if __name__ == "__main__":
    main()
