"""GPU (-m gpu): the full AutoMoE forward through the drop-in module against the oracle and
the reference's golden vectors.  fp32 mode: 1e-4 relative; bf16 mode: 1e-2 relative
(BASELINE.md §5); top-1 routing identical."""
import numpy as np
import pytest
import torch

from _util import build_b200_model, golden_batch, golden_config, rel_err, rel_l2
from oracle import automoe_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SMALL = ["waypoints", "speed", "speed_seq", "expert_weights", "context_features", "combined_features", "gate_logits"]


@pytest.fixture(scope="module")
def model_sd():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return build_b200_model(DEV, "auto")


def _to(batch, dev):
    return {k: v.to(dev) for k, v in batch.items()}


@pytest.mark.parametrize("name", ["automoe_b2_64", "automoe_b1_256", "automoe_b3_96_seq"])
def test_fp32_matches_reference_golden(name, golden_dir, model_sd):
    m, _ = model_sd
    g = np.load(golden_dir / f"{name}.npz")
    batch = _to(golden_batch(g), DEV)
    with torch.no_grad():
        out = m(batch)  # no autocast -> fp32 kernels
    sub = int(g["sub"])
    gk = {"speed_seq": "speed_seq_out"}
    for k in SMALL:
        assert rel_err(out[k].cpu(), g[gk.get(k, k)]) < 1e-4, (k, rel_err(out[k].cpu(), g[gk.get(k, k)]))
    det = out["expert_outputs"][0]
    assert rel_err(det["class_logits"].cpu(), g["det_class_logits"]) < 1e-4
    assert rel_err(det["bbox_deltas"].cpu(), g["det_bbox_deltas"]) < 1e-4
    assert rel_err(out["expert_outputs"][1][:, :, ::sub, ::sub].cpu(), g["seg_sub"]) < 1e-4
    assert rel_err(out["expert_outputs"][2][:, :, ::sub, ::sub].cpu(), g["drv_sub"]) < 1e-4
    assert rel_err(out["expert_outputs"][1].mean(dim=(2, 3)).cpu(), g["seg_mean"]) < 1e-4
    assert np.array_equal(out["expert_weights"].argmax(1).cpu().numpy(), g["expert_weights"].argmax(1))
    assert out["expert_outputs"][1].shape == (int(g["B"]), 19, int(g["H"]), int(g["W"]))
    assert rel_err(m.get_expert_weights(batch).cpu(), g["ctx_only_weights"]) < 1e-4


def test_sigmoid_gate_matches_reference_golden(golden_dir):
    """use_softmax=False (gating_network.py:159-160), reachable through AutoMoE's gating config: fp32 vs the reference's
    golden outputs, bf16 (tensor-core gate variant at B >= 16) vs the oracle, and the training combine's gradients."""
    g = np.load(golden_dir / "automoe_b2_64_sigmoid.npz")
    cfg = golden_config(g)
    m, sd = build_b200_model(DEV, "auto", config=cfg)
    batch = _to(golden_batch(g), DEV)
    with torch.no_grad():
        out = m(batch)
    for k in SMALL:
        assert rel_err(out[k].cpu(), g[{"speed_seq": "speed_seq_out"}.get(k, k)]) < 1e-4, k
    assert rel_err(m.get_expert_weights(batch).cpu(), g["ctx_only_weights"]) < 1e-4
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    big = _to(synth.synth_batch(32, 64, 64, seed=9), DEV)
    with torch.no_grad():
        ref = O.automoe_forward(sdd, big, cfg)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o16 = m(big)
    assert rel_err(o16["expert_weights"], ref["expert_weights"]) < 1e-2
    assert rel_err(o16["combined_features"], ref["combined_features"]) < 2e-2
    # differentiable combine (training path): gradients of the sigmoid gate against torch autograd
    from automoe_b200.training import functional as TF
    gen = torch.Generator().manual_seed(4)
    lg = torch.randn((6, 3), generator=gen).to(DEV).requires_grad_(True)
    pr = [torch.randn((6, 256), generator=gen).to(DEV).requires_grad_(True) for _ in range(3)]
    w, c = TF.gate_combine(lg, pr, 1.0, use_softmax=False)
    lg2 = lg.detach().clone().requires_grad_(True)
    pr2 = [p.detach().clone().requires_grad_(True) for p in pr]
    w2 = torch.sigmoid(lg2)
    w2 = w2 / (w2.sum(dim=1, keepdim=True) + 1e-8)
    c2 = sum(w2[:, i:i + 1] * pr2[i] for i in range(3))
    gw, gc = torch.randn_like(w), torch.randn_like(c)
    ((w * gw).sum() + (c * gc).sum()).backward()
    ((w2 * gw).sum() + (c2 * gc).sum()).backward()
    assert rel_err(w, w2) < 1e-6 and rel_err(c, c2) < 1e-6
    assert rel_err(lg.grad, lg2.grad) < 1e-5
    for a, b in zip(pr, pr2):
        assert rel_err(a.grad, b.grad) < 1e-6


def test_four_expert_shipped_config_matches_reference_golden(golden_dir):
    """The shipped model_config.json (3 BDD experts + image-only nuScenes expert, models/experts/nuscenes_expert.py:96-190 +
    NuScenesExpertExtractor): fp32 against the reference's golden outputs, bf16 against the oracle, graph capture."""
    g = np.load(golden_dir / "automoe4_b2_96.npz")
    cfg = golden_config(g)
    assert len(cfg["experts"]) == 4
    m, sd = build_b200_model(DEV, "auto", config=cfg)
    batch = _to(golden_batch(g), DEV)
    with torch.no_grad():
        out = m(batch)
    for k in SMALL:
        assert rel_err(out[k].cpu(), g[{"speed_seq": "speed_seq_out"}.get(k, k)]) < 1e-4, (k, rel_err(out[k].cpu(), g[{"speed_seq": "speed_seq_out"}.get(k, k)]))
    assert out["expert_weights"].shape == (2, 4)
    nus = out["expert_outputs"][3]
    assert set(nus.keys()) == {"class_logits", "bbox_preds"} and nus["class_logits"].shape == (2, 196, 10) and nus["bbox_preds"].shape == (2, 196, 4)
    assert rel_err(nus["class_logits"].cpu(), g["nus_class_logits"]) < 1e-4
    assert rel_err(nus["bbox_preds"].cpu(), g["nus_bbox_preds"]) < 1e-4
    assert rel_err(out["expert_outputs"][0]["class_logits"].cpu(), g["det_class_logits"]) < 1e-4
    assert np.array_equal(out["expert_weights"].argmax(1).cpu().numpy(), g["expert_weights"].argmax(1))
    assert rel_err(m.get_expert_weights(batch).cpu(), g["ctx_only_weights"]) < 1e-4
    # stand-alone expert + extractor calls, as the reference's unit tests make them
    with torch.no_grad():
        alone = m.experts[3]({"image": batch["image"]})
        feat = m.expert_extractors.extractors[3](alone)
    assert rel_err(alone["class_logits"], nus["class_logits"]) < 1e-6 and feat.shape == (2, 256)
    # bf16 at a tensor-core batch + graph replay
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    big = _to(synth.synth_batch(32, 128, 128, seed=9), DEV)
    with torch.no_grad():
        ref = O.automoe_forward(sdd, big, cfg)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o16 = m(big)
            ref16 = O.automoe_forward(sdd, big, cfg)
    for k in SMALL:
        ours, theirs = rel_err(o16[k], ref[k]), rel_err(ref16[k].float(), ref[k])
        assert ours < max(1e-2, 1.25 * theirs), (k, ours, theirs)
    ours, theirs = rel_err(o16["expert_outputs"][3]["class_logits"], ref["expert_outputs"][3]["class_logits"]), \
        rel_err(ref16["expert_outputs"][3]["class_logits"].float(), ref["expert_outputs"][3]["class_logits"])
    assert ours < max(1e-2, 1.25 * theirs), (ours, theirs)
    gr = m.capture(big)
    rep = gr()
    torch.cuda.synchronize()
    for k in SMALL:
        assert torch.equal(rep[k], o16[k]), k


def test_fp32_matches_oracle_batch(model_sd):
    """Bigger seeded batch against the oracle run on the same device (TF32 off)."""
    m, sd = model_sd
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    batch = _to(synth.synth_batch(8, 128, 128, seed=3), DEV)
    with torch.no_grad():
        out = m(batch)
        ref = O.automoe_forward(sdd, batch, synth.CONFIG_3EXPERT)
    for k in SMALL:
        assert rel_err(out[k], ref[k]) < 1e-4, (k, rel_err(out[k], ref[k]))
    assert rel_err(out["expert_outputs"][1], ref["expert_outputs"][1]) < 1e-4
    assert rel_err(out["expert_outputs"][2], ref["expert_outputs"][2]) < 1e-4
    assert torch.equal(out["expert_weights"].argmax(1), ref["expert_weights"].argmax(1))


@pytest.mark.parametrize("B,H", [(4, 64), (2, 256), (16, 256)])
def test_bf16_within_tolerance(model_sd, B, H):
    """bf16 tcgen05 path vs the fp32 oracle and vs the reference's own bf16-autocast arithmetic."""
    m, sd = model_sd
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    batch = _to(synth.synth_batch(B, H, H, seed=5), DEV)
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(batch)
            ref_bf16 = O.automoe_forward(sdd, batch, synth.CONFIG_3EXPERT)
        ref = O.automoe_forward(sdd, batch, synth.CONFIG_3EXPERT)
    report = {}
    for k in SMALL:
        ours, theirs = rel_err(out[k], ref[k]), rel_err(ref_bf16[k].float(), ref[k])
        report[k] = (round(ours, 5), round(theirs, 5))
        # within 1e-2 relative, or at least as accurate as the reference's own bf16 path
        assert ours < max(1e-2, 1.25 * theirs), (k, ours, theirs)
    for i in (1, 2):
        ours = rel_err(out["expert_outputs"][i].float(), ref["expert_outputs"][i])
        theirs = rel_err(ref_bf16["expert_outputs"][i].float(), ref["expert_outputs"][i])
        report[f"expert{i}"] = (round(ours, 5), round(theirs, 5))
        assert ours < max(1e-2, 1.25 * theirs), (i, ours, theirs)
        assert out["expert_outputs"][i].dtype == torch.bfloat16
    print("bf16 rel err (ours, reference-autocast):", report)
    # routing: identical top-1 wherever the fp32 logit gap exceeds the bf16 noise floor
    top2 = ref["gate_logits"].topk(2, dim=1).values
    gap = top2[:, 0] - top2[:, 1]
    noise = (out["gate_logits"] - ref["gate_logits"]).abs().max().item()
    safe = gap > 2 * noise
    assert torch.equal(out["expert_weights"].argmax(1)[safe], ref["expert_weights"].argmax(1)[safe])


def test_experts_standalone_and_batch1(model_sd):
    """Experts are callable on their own (expert trainers / evals do that) and give the grouped result."""
    m, sd = model_sd
    batch = _to(synth.synth_batch(1, 64, 64, seed=11), DEV)
    with torch.no_grad():
        full = m(batch)
        seg = m.experts[1](batch["image"])
        det = m.experts[0](batch["image"])
    assert rel_err(seg, full["expert_outputs"][1]) < 1e-6
    assert rel_err(det["class_logits"], full["expert_outputs"][0]["class_logits"]) < 1e-6


def test_weight_update_invalidates_pack(model_sd):
    m, sd = model_sd
    batch = _to(synth.synth_batch(2, 64, 64, seed=12), DEV)
    with torch.no_grad():
        a = m(batch)["waypoints"].clone()
        m.policy_head.head_wp[4].bias.add_(1.0)
        b = m(batch)["waypoints"]
        m.policy_head.head_wp[4].bias.sub_(1.0)
    assert torch.allclose(b - a, torch.ones_like(a), atol=1e-5)


def test_l2_chunked_stem_layer1_is_bit_identical(model_sd, monkeypatch):
    """Walking stem+layer1 in sub-batches (L2-resident chunks, ragged last chunk, strided write into the
    full grouped tensor) must not change a single bit of the forward."""
    m, sd = model_sd
    batch = _to(synth.synth_batch(16, 256, 256, seed=7), DEV)
    outs = {}
    for chunk in ("0", "6"):
        monkeypatch.setenv("AMOE_L2_CHUNK", chunk)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            outs[chunk] = m(batch)
    for k in SMALL:
        assert torch.equal(outs["0"][k], outs["6"][k]), k
    for i in (1, 2):
        assert torch.equal(outs["0"]["expert_outputs"][i], outs["6"]["expert_outputs"][i])
    for k in ("class_logits", "bbox_deltas"):
        assert torch.equal(outs["0"]["expert_outputs"][0][k], outs["6"]["expert_outputs"][0][k])


def test_fp32_mode_tensor_core_and_cuda_core_paths_agree(model_sd, monkeypatch):
    """fp32 (parity) mode: the split-operand tensor-core convolutions (AMOE_F32_TC=1, default) against the CUDA-core fp32
    kernels on the same batch - both are fp32-accurate, so they agree to a few 1e-6 and give the same routing."""
    m, sd = model_sd
    batch = _to(synth.synth_batch(6, 128, 128, seed=23), DEV)
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("AMOE_F32_TC", flag)
        with torch.no_grad():
            outs[flag] = m(batch)
    for k in SMALL:
        assert rel_err(outs["1"][k], outs["0"][k]) < 2e-5, (k, rel_err(outs["1"][k], outs["0"][k]))
    assert rel_err(outs["1"]["expert_outputs"][1], outs["0"]["expert_outputs"][1]) < 2e-5
    assert torch.equal(outs["1"]["expert_weights"].argmax(1), outs["0"]["expert_weights"].argmax(1))


def test_dual_stage_entry_launch_is_bit_identical(model_sd, monkeypatch):
    """Stage-entry conv1 (3x3/s2) + the block's 1x1/s2 downsample as ONE dual-problem launch of the tcgen05 kernel
    (amoe_conv2d_dual_fwd) against the two separate launches: same MMAs in the same order -> the same bits, for the padded
    (layer2, 256x256 frames), the unpadded and the odd-sized geometries."""
    m, sd = model_sd
    for B, H in ((16, 256), (3, 96), (2, 160)):
        batch = _to(synth.synth_batch(B, H, H, seed=17), DEV)
        outs = {}
        for flag in ("0", "1"):
            monkeypatch.setenv("AMOE_DUAL", flag)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                outs[flag] = m(batch)
        for k in SMALL:
            assert torch.equal(outs["0"][k], outs["1"][k]), (B, H, k)
        for i in (1, 2):
            assert torch.equal(outs["0"]["expert_outputs"][i], outs["1"]["expert_outputs"][i]), (B, H, i)
        assert torch.equal(outs["0"]["expert_outputs"][0]["class_logits"], outs["1"]["expert_outputs"][0]["class_logits"])


def test_policy_backbone_on_second_stream_is_bit_identical(model_sd, monkeypatch):
    """The policy backbone forked onto its own stream behind the experts' last tensor-bound launch (AMOE_TAIL_OVERLAP, default
    on) runs the same kernels on the same data as the single-stream order: identical bits, eager and through a CUDA graph."""
    m, sd = model_sd
    for B, H in ((16, 256), (3, 96)):
        batch = _to(synth.synth_batch(B, H, H, seed=23), DEV)
        outs = {}
        for flag in ("0", "1"):
            monkeypatch.setenv("AMOE_TAIL_OVERLAP", flag)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                outs[flag] = m(batch)
            torch.cuda.synchronize()
        for k in SMALL:
            assert torch.equal(outs["0"][k], outs["1"][k]), (B, H, k)
        for i in (1, 2):
            assert torch.equal(outs["0"]["expert_outputs"][i], outs["1"]["expert_outputs"][i]), (B, H, i)
    monkeypatch.setenv("AMOE_TAIL_OVERLAP", "1")
    g = m.capture(batch)
    out = g(batch)
    torch.cuda.synchronize()
    for k in SMALL:
        assert torch.equal(out[k], outs["0"][k]), k


def test_cuda_graph_capture_replays_the_eager_forward(model_sd):
    """AutoMoE.capture(): replaying the captured graph on new inputs gives bit-identical outputs to the eager
    call (same kernels, same order; side-stream logit writers joined inside the graph)."""
    m, sd = model_sd
    b1 = _to(synth.synth_batch(8, 256, 256, seed=31), DEV)
    b2 = _to(synth.synth_batch(8, 256, 256, seed=32), DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        eager2 = m(b2)
    g = m.capture(b1)
    assert g.launches_per_replay >= 25
    out = g(b2)
    torch.cuda.synchronize()
    for k in SMALL:
        assert torch.equal(out[k], eager2[k]), k
    for i in (1, 2):
        assert torch.equal(out["expert_outputs"][i], eager2["expert_outputs"][i])
    out1 = {k: v.clone() for k, v in g(b1).items() if torch.is_tensor(v)}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        eager1 = m(b1)
    for k in SMALL:
        assert torch.equal(out1[k], eager1[k]), k
