"""GPU (-m gpu, needs >= 2 devices): every C-ABI entry runs on its context's device, whatever the calling thread's
current device is (a model built with create_automoe_model(cfg, 'cuda:1') while the current device is 0)."""
import pytest
import torch

from _util import build_b200_model, rel_err
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_cuda1_while_current_device_is_0():
    torch.cuda.set_device(0)
    m0, _ = build_b200_model("cuda:0", "auto")
    m1, _ = build_b200_model("cuda:1", "auto")
    batch = synth.synth_batch(4, 64, 64, seed=3)
    b0 = {k: v.to("cuda:0") for k, v in batch.items()}
    b1 = {k: v.to("cuda:1") for k, v in batch.items()}
    assert torch.cuda.current_device() == 0
    with torch.no_grad():
        o0 = m0(b0)
        o1 = m1(b1)                      # current device is still 0
        with torch.autocast("cuda", dtype=torch.bfloat16):
            p0, p1 = m0(b0), m1(b1)
    assert torch.cuda.current_device() == 0
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    for k in ("waypoints", "speed", "expert_weights", "gate_logits"):
        assert o1[k].device.index == 1
        assert torch.equal(o0[k].cpu(), o1[k].cpu()), k
        assert torch.equal(p0[k].cpu(), p1[k].cpu()), k
    assert rel_err(o1["expert_outputs"][1].cpu(), o0["expert_outputs"][1].cpu()) == 0.0
