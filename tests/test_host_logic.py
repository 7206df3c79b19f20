"""CPU: host-side logic of the drop-in modules (no kernels run)."""
import numpy as np
import pytest
import torch

from oracle import automoe_oracle as O
from oracle import synth


@pytest.fixture(scope="module")
def model():
    from automoe_b200.models.automoe import create_automoe_model
    return create_automoe_model(synth.CONFIG_3EXPERT, "cpu").eval()


def test_state_dict_keys_match_reference(model, golden_dir):
    keys = (golden_dir / "state_dict_keys.txt").read_text().split()
    assert list(model.state_dict().keys()) == keys
    assert len(keys) == 466
    assert sum(p.numel() for p in model.parameters()) == 39_949_157


def test_api_surface(model):
    for attr in ("experts", "expert_extractors", "context_extractor", "gating_network", "policy_head",
                 "forward", "get_expert_weights", "load_expert_checkpoints", "freeze_experts", "unfreeze_experts"):
        assert hasattr(model, attr)
    model.freeze_experts()
    assert all(not p.requires_grad for e in model.experts for p in e.parameters())
    assert all(p.requires_grad for p in model.gating_network.parameters())
    model.unfreeze_experts()
    assert all(p.requires_grad for e in model.experts for p in e.parameters())
    trainable = sum(p.numel() for n, p in model.named_parameters() if not n.startswith("experts."))
    assert trainable == 2_870_657  # extractors + context + gating + policy


@pytest.mark.parametrize("speed_seq,controls", [(1, "col"), (5, "seq"), (3, "missing"), (4, "3d")])
def test_vehicle_state_matches_oracle(model, speed_seq, controls):
    B = 3
    g = torch.Generator().manual_seed(0)
    batch = {"speed": torch.rand(B, speed_seq, generator=g)}
    if controls == "col":
        for k in ("steering", "throttle", "brake"):
            batch[k] = torch.rand(B, 1, generator=g)
    elif controls == "seq":
        for k in ("steering", "throttle", "brake"):
            batch[k] = torch.rand(B, speed_seq, generator=g)
    elif controls == "3d":
        for k in ("steering", "throttle", "brake"):
            batch[k] = torch.rand(B, 2, 2, generator=g)
    assert torch.equal(model._vehicle_state(batch), O.vehicle_state(batch))


def test_gate_param_layout_size(model):
    """flat buffer length == what csrc/gate.cu derives from the dims: every tensor starts on an 8-element
    boundary, so it is 32-byte aligned in fp32 and 16-byte aligned in the bf16 copy the tensor-core variant reads."""
    from automoe_b200 import _ops
    from automoe_b200.models._gatepack import gate_param_tensors
    n_ch = [14, 19, 3]
    ts = gate_param_tensors(model.context_extractor, list(model.expert_extractors.extractors), model.gating_network,
                            n_ch, 64, 128)
    al8 = lambda n: (n + 7) & ~7
    total = sum(al8(t.numel()) for t in ts)
    raw = sum(t.numel() for t in ts)
    assert raw == 2_400 + 415_488 + 602_115  # context + extractors + gating
    assert total >= raw and total - raw < 8 * len(ts)
    flat = _ops.flat_params(ts, "cpu")
    assert flat.numel() == total and flat.dtype == torch.float32
    off = 0
    for t in ts:                                      # tensors sit at their aligned offsets, padding is zero
        assert off % 8 == 0
        assert torch.equal(flat[off:off + t.numel()], t.detach().float().reshape(-1))
        assert (flat[off + t.numel():off + al8(t.numel())] == 0).all()
        off += al8(t.numel())
    assert _ops.mlp_tc(torch.bfloat16) and not _ops.mlp_tc(torch.float32)   # fp32 mode never takes the TF32 kernels


def test_train_mode_is_rejected(model):
    model.train()
    try:
        with pytest.raises(NotImplementedError):
            model({"image": torch.zeros(1, 3, 64, 64), "speed": torch.zeros(1, 1)})
    finally:
        model.eval()


def test_unsupported_configs_fail_loudly():
    from automoe_b200.models.automoe import create_automoe_model
    cfg = {k: v for k, v in synth.CONFIG_3EXPERT.items()}
    cfg["experts"] = list(cfg["experts"]) + [{"type": "nuscenes", "use_lidar": True, "pretrained_backbone": False}]
    with pytest.raises(NotImplementedError, match="use_lidar"):      # PointNet branch: outside the hot path, named in the error
        create_automoe_model(cfg, "cpu")
    cfg = dict(synth.CONFIG_3EXPERT)
    cfg["context"] = {"type": "full"}
    with pytest.raises(NotImplementedError):
        create_automoe_model(cfg, "cpu")


def test_shipped_four_expert_config_builds_with_reference_keys():
    """models/configs/automoe/model_config.json (3 BDD experts + image-only nuScenes expert, sigmoid/softmax gate keys that
    AutoMoE does not forward) builds; nuScenes sub-modules carry the reference's parameter names and shapes."""
    from automoe_b200.models.automoe import create_automoe_model
    m = create_automoe_model(synth.CONFIG_4EXPERT, "cpu")
    sd = m.state_dict()
    assert len(sd) == 609
    assert sd["experts.3.image_backbone.7.1.bn2.running_var"].shape == (512,)
    assert sd["experts.3.query_embed.weight"].shape == (196, 256)
    assert sd["experts.3.bbox_head.weight"].shape == (4, 128)
    assert sd["expert_extractors.extractors.3.feature_extractor.0.weight"].shape == (512, 196 * 14)
    assert sd["gating_network.gate_network.0.weight"].shape == (128, 128 + 4 * 256)



def test_matcher_pack_targets_is_one_padded_pack():
    """Ragged targets -> padded [B,Nmax,D] / [B,Nmax] (-1) without per-image copies; empty images stay padding."""
    import torch
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    g = torch.Generator().manual_seed(0)
    counts = [3, 0, 5, 1]
    targets = [{"boxes": torch.rand(n, 4, generator=g), "labels": torch.randint(0, 10, (n,), generator=g)} for n in counts]
    tb, tl = HungarianMatcher.pack_targets(targets, counts, 5, 4, "cpu")
    for b, t in enumerate(targets):
        n = counts[b]
        assert torch.equal(tb[b, :n], t["boxes"]) and torch.equal(tl[b, :n], t["labels"])
        assert (tl[b, n:] == -1).all() and (tb[b, n:] == 0).all()


def test_stem_mode_follows_geometry(monkeypatch):
    """bf16: the tensor-core stem for frames it takes (even H/W, W <= 256), the row-window variant otherwise
    (BDD's native 720x1280) - chosen by the module, no environment switch needed; fp32 always CUDA cores."""
    import torch
    from automoe_b200 import _ops
    monkeypatch.delenv("AMOE_STEM", raising=False)
    assert _ops.stem_mode(torch.bfloat16, 256, 256) == "tc"
    assert _ops.stem_mode(torch.bfloat16, 720, 1280) == "rowwin"
    assert _ops.stem_mode(torch.bfloat16, 255, 256) == "rowwin"
    assert _ops.stem_mode(torch.float32, 256, 256) == "simt"
    monkeypatch.setenv("AMOE_STEM", "rowwin")
    assert _ops.stem_mode(torch.bfloat16, 256, 256) == "rowwin"


def test_grouped_train_forward_selection(model):
    """Which frozen-expert train-mode forwards take the lockstep (grouped) path: same-geometry frozen ResNet-18 experts in
    train mode on frames whose stage inputs stay even; anything else runs expert by expert (host logic only)."""
    import copy
    from automoe_b200.models.experts._trunk import grouped_train_supported
    m = copy.deepcopy(model)
    m.freeze_experts()
    m.train()
    ex = m._bdd_experts()
    assert grouped_train_supported(ex, torch.zeros(2, 3, 256, 256))
    assert grouped_train_supported(ex, torch.zeros(2, 3, 96, 160))
    assert not grouped_train_supported(ex, torch.zeros(2, 3, 90, 122))        # odd stage sizes
    assert not grouped_train_supported(ex, torch.zeros(2, 3, 720, 1280))      # 720 / 16 = 45 rows into layer4
    assert not grouped_train_supported(ex[:1], torch.zeros(2, 3, 256, 256))   # one expert: nothing to group
    ex[1].backbone[5][0].bn1.eval()
    assert not grouped_train_supported(ex, torch.zeros(2, 3, 256, 256))       # a BatchNorm on running statistics
    ex[1].backbone[5][0].bn1.train()
    ex[0].backbone[6][0].downsample[1].eval()
    assert not grouped_train_supported(ex, torch.zeros(2, 3, 256, 256))       # ... also a downsample BatchNorm
    ex[0].backbone[6][0].downsample[1].train()
    ex[2].backbone[7][1].bn2.momentum = None
    assert not grouped_train_supported(ex, torch.zeros(2, 3, 256, 256))       # cumulative average: not what the grouped pass computes
    ex[2].backbone[7][1].bn2.momentum = 0.1
    assert grouped_train_supported(ex, torch.zeros(2, 3, 256, 256))
    next(ex[2].parameters()).requires_grad_(True)
    assert not grouped_train_supported(ex, torch.zeros(2, 3, 256, 256))       # an expert that trains needs the autograd path


def test_graphed_train_step_and_device_counters_need_cuda():
    """No CPU fallback: the flat optimizer (and with it the graphed step) refuses CPU parameters."""
    from automoe_b200.training.train_gating_network import FlatAdamW
    lin = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FlatAdamW(lin.parameters())


def test_even_size_rule_for_stride2_training_convs():
    """Stride-2 training convolutions on odd sizes run on the tensor cores with one zero row / column appended - only where the
    output size stays what nn.Conv2d gives (the extra row then stands where the zero padding is)."""
    import torch.nn.functional as F
    from automoe_b200.training.functional import _even_size
    for (H, W, K, p) in [(45, 80, 3, 1), (45, 80, 1, 0), (13, 11, 3, 1), (45, 81, 7, 3), (9, 9, 5, 2), (44, 80, 3, 1)]:
        Ho, Wo = (H + 2 * p - K) // 2 + 1, (W + 2 * p - K) // 2 + 1
        Hk, Wk = _even_size(H, W, K, K, 2, p, Ho, Wo)
        assert Hk % 2 == 0 and Wk % 2 == 0 and Hk - H in (0, 1) and Wk - W in (0, 1)
        x = torch.randn(1, 2, H, W)
        w = torch.randn(3, 2, K, K)
        ref = F.conv2d(x, w, None, 2, p)
        padded = F.conv2d(F.pad(x, (0, Wk - W, 0, Hk - H)), w, None, 2, p)
        assert ref.shape == padded.shape == (1, 3, Ho, Wo) and torch.allclose(ref, padded, atol=1e-5)
    assert _even_size(45, 80, 3, 3, 1, 1, 45, 80) == (45, 80)           # stride 1: untouched
    assert _even_size(45, 80, 2, 2, 2, 0, 22, 40) == (45, 80)           # 2x2/s2/p0: one more row would add an output row
