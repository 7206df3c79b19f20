"""Drop-in proof at the reference's own call sites.

CPU part (runs where /root/reference exists, i.e. in the build container; skipped elsewhere): the SOURCE of the
reference's `load_model` and `model_infer` (inference/run_automoe.py:34-53,144-156) is extracted with `ast` - nothing of
it is copied into this repository - and executed with `create_automoe_model` bound to THIS package's factory:
  * load_model builds the module from a config JSON, loads a DDP-prefixed checkpoint produced from the reference's key
    list, and leaves it in eval mode with every tensor in place (0 missing / 0 unexpected keys);
  * model_infer builds its batch from a uint8 camera frame with the reference's own torchvision transform and calls
    model(batch) under the reference's autocast; on a CPU-only box the call must get as far as this package's device
    check (the module has no CPU path by design) - everything the reference does before the kernels is accepted.
GPU part (-m gpu): the same call sequence re-enacted with torchvision on the GPU box against the device-side mirror
(automoe_b200.inference.model_infer), including the reference's default float16 autocast.
"""
import ast
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import synth

REF = Path("/root/reference/inference/run_automoe.py")


def _reference_functions(names):
    tree = ast.parse(REF.read_text())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names)
    from typing import Any, Dict, Optional, Tuple
    import torch.nn as nn
    from torchvision import transforms as T
    from automoe_b200.models.automoe import create_automoe_model
    ns = {"json": json, "Path": Path, "torch": torch, "nn": nn, "np": np, "T": T, "Dict": Dict, "Any": Any, "Optional": Optional,
          "Tuple": Tuple, "create_automoe_model": create_automoe_model}
    exec(compile(ast.Module(body=body, type_ignores=[]), str(REF), "exec"), ns)
    return [ns[n] for n in names]


@pytest.mark.skipif(not REF.exists(), reason="/root/reference is only present in the build container")
def test_reference_load_model_and_model_infer_run_over_this_package(tmp_path, golden_dir):
    load_model, model_infer, build_image_transform = _reference_functions(["load_model", "model_infer", "build_image_transform"])
    from automoe_b200.models.automoe import AutoMoE, create_automoe_model
    cfg = dict(synth.CONFIG_3EXPERT)
    (tmp_path / "model_config.json").write_text(json.dumps(cfg))
    keys = (golden_dir / "state_dict_keys.txt").read_text().split()
    template = create_automoe_model(cfg, "cpu").state_dict()
    assert list(template.keys()) == keys                      # the reference's 466 keys, in order
    sd = synth.synth_state_dict(template, 0)
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}, "epoch": 3}, tmp_path / "ckpt.pt")
    model = load_model(str(tmp_path / "model_config.json"), str(tmp_path / "ckpt.pt"), torch.device("cpu"))
    assert isinstance(model, AutoMoE) and not model.training
    got = model.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in keys)
    # model_infer: the reference's own batch construction and autocast around this package's forward
    frame = synth.synth_u8_frame(600, 800, 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model_infer(model, frame, 12.0, torch.device("cpu"), build_image_transform((256, 256)))


@pytest.mark.gpu
def test_reference_call_sequence_on_gpu_equals_device_mirror():
    """What model_infer does, step by step, with the real torchvision/Pillow transform on the host and the reference's
    default autocast (float16 requested -> this package computes bf16 and says so), against the device-side mirror."""
    T = pytest.importorskip("torchvision.transforms")
    from automoe_b200.inference import build_image_transform, model_infer
    from _util import build_b200_model
    dev = torch.device("cuda:0")
    m, _ = build_b200_model("cuda:0", "auto")
    frame = synth.synth_u8_frame(600, 800, 6)
    tf = T.Compose([T.ToPILImage(), T.Resize((256, 256), interpolation=T.InterpolationMode.BILINEAR), T.ToTensor(),
                    T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    tensor = tf(frame).unsqueeze(0).to(dev)
    batch = {"image": tensor, "speed": torch.tensor([[17.0]], device=dev), "steering": torch.zeros(1, 1, device=dev),
             "throttle": torch.zeros(1, 1, device=dev), "brake": torch.zeros(1, 1, device=dev)}
    with torch.no_grad(), pytest.warns(UserWarning, match="float16"):
        import automoe_b200.models._precision as P
        P._warned_fp16 = False
        with torch.autocast(device_type="cuda", enabled=True):
            ref_style = m(batch)
    mirror = model_infer(m, frame, 17.0, dev, build_image_transform((256, 256)))
    for k in ("waypoints", "speed", "speed_seq", "expert_weights", "gate_logits"):
        assert torch.equal(ref_style[k], mirror[k]), k
    assert torch.equal(ref_style["expert_outputs"][1], mirror["expert_outputs"][1])
