"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(imported from /root/reference; CPU, fp32) on seeded synthetic weights and inputs.

    python tests/golden/make_golden.py

Needs /root/reference (only present in the build container); the produced .npz files are
committed.  Weights and inputs are NOT stored: they are regenerated at test time from the
same seeds by oracle/synth.py.  Versions used to generate are recorded in each file.
"""
import sys
from pathlib import Path

import numpy as np
import scipy
import torch
import torchvision

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from models.automoe import create_automoe_model as ref_create  # noqa: E402  (reference)
from training.hungarian_matcher import HungarianMatcher as RefMatcher  # noqa: E402  (reference)

from oracle import synth  # noqa: E402

OUT = Path(__file__).resolve().parent
VERS = np.array([f"torch {torch.__version__}", f"torchvision {torchvision.__version__}", f"scipy {scipy.__version__}"])


def automoe_case(name, B, H, W, sub, speed_seq=1, controls_seq=False, use_softmax=True, four_experts=False):
    torch.manual_seed(0)
    cfg = dict(synth.CONFIG_4EXPERT if four_experts else synth.CONFIG_3EXPERT)
    cfg["gating"] = dict(cfg["gating"], use_softmax=use_softmax)
    if four_experts:
        # NuScenesExpert.__init__ always asks torchvision for ImageNet weights (nuscenes_expert.py:108); there is no network
        # here and every weight is overwritten by the synthetic state_dict below, so resnet18 is asked for random weights
        import torchvision.models as tvm
        orig = tvm.resnet18
        tvm.resnet18 = lambda pretrained=False, **k: orig(weights=None)
    ref = ref_create(cfg, "cpu").eval()
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), 0), strict=True)
    batch = synth.synth_batch(B, H, W, seed=1, speed_seq=speed_seq)
    if controls_seq:
        g = torch.Generator().manual_seed(7)
        for k in ("steering", "throttle", "brake"):
            batch[k] = torch.rand((B, speed_seq), generator=g) - 0.5
    with torch.no_grad():
        r = ref(batch)
        w_ctx = ref.get_expert_weights(batch)
    seg, drv = r["expert_outputs"][1], r["expert_outputs"][2]
    np.savez_compressed(
        OUT / f"{name}.npz", versions=VERS, B=B, H=H, W=W, sub=sub, speed_seq=speed_seq, controls_seq=controls_seq,
        waypoints=r["waypoints"].numpy(), speed=r["speed"].numpy(), speed_seq_out=r["speed_seq"].numpy(),
        expert_weights=r["expert_weights"].numpy(), context_features=r["context_features"].numpy(),
        combined_features=r["combined_features"].numpy(), gate_logits=r["gate_logits"].numpy(),
        det_class_logits=r["expert_outputs"][0]["class_logits"].numpy(),
        det_bbox_deltas=r["expert_outputs"][0]["bbox_deltas"].numpy(),
        seg_sub=seg[:, :, ::sub, ::sub].numpy(), drv_sub=drv[:, :, ::sub, ::sub].numpy(),
        seg_mean=seg.mean(dim=(2, 3)).numpy(), drv_mean=drv.mean(dim=(2, 3)).numpy(),
        seg_abs_sum=np.float64(seg.double().abs().sum().item()), drv_abs_sum=np.float64(drv.double().abs().sum().item()),
        ctx_only_weights=w_ctx.numpy(), use_softmax=use_softmax, four_experts=four_experts,
        **({"nus_class_logits": r["expert_outputs"][3]["class_logits"].numpy(),
            "nus_bbox_preds": r["expert_outputs"][3]["bbox_preds"].numpy()} if four_experts else {}),
    )
    print(name, "weights", r["expert_weights"].numpy().round(4).tolist())


def matcher_case(name, B, Q, C, D, n_min, n_max, seed):
    outputs, targets = synth.synth_matcher_case(B, Q, C, D, n_min, n_max, seed)
    m = RefMatcher()
    idx = m(outputs, targets)
    # per-image reference cost matrices, recomputed with the reference's own expression
    from torchvision.ops import box_convert, generalized_box_iou
    costs = []
    for b in range(B):
        prob = outputs["pred_logits"][b].softmax(-1)
        tl, tb, pb = targets[b]["labels"], targets[b]["boxes"], outputs["pred_boxes"][b]
        cc = -prob[:, tl]
        cb = torch.cdist(pb, tb, p=1)
        if D == 4:
            cg = -generalized_box_iou(box_convert(pb, "cxcywh", "xyxy"), box_convert(tb, "cxcywh", "xyxy"))
        elif D == 7:
            def bev(x):
                return torch.stack([x[:, 0] - x[:, 3] / 2, x[:, 1] - x[:, 4] / 2, x[:, 0] + x[:, 3] / 2, x[:, 1] + x[:, 4] / 2], 1)
            cg = -generalized_box_iou(bev(pb), bev(tb))
        else:
            cg = torch.zeros_like(cb)
        costs.append((5.0 * cb + 1.0 * cc + 2.0 * cg).numpy())
    d = dict(versions=VERS, B=B, Q=Q, C=C, D=D, n_min=n_min, n_max=n_max, seed=seed)
    for b in range(B):
        d[f"rows_{b}"] = idx[b][0].numpy()
        d[f"cols_{b}"] = idx[b][1].numpy()
        d[f"cost_{b}"] = costs[b]
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, [len(i[0]) for i in idx])


def state_dict_keys():
    ref = ref_create(synth.CONFIG_3EXPERT, "cpu")
    (OUT / "state_dict_keys.txt").write_text("\n".join(ref.state_dict().keys()) + "\n")


if __name__ == "__main__":
    state_dict_keys()
    automoe_case("automoe_b2_64", B=2, H=64, W=64, sub=4)
    automoe_case("automoe_b1_256", B=1, H=256, W=256, sub=16)                 # BASELINE.json configs[0]
    automoe_case("automoe_b3_96_seq", B=3, H=96, W=96, sub=8, speed_seq=5, controls_seq=True)
    automoe_case("automoe_b2_64_sigmoid", B=2, H=64, W=64, sub=4, use_softmax=False)     # gating_network.py:159-160
    automoe_case("automoe4_b2_96", B=2, H=96, W=96, sub=8, four_experts=True)            # shipped model_config.json (4 experts)
    matcher_case("matcher_d4_q64", B=4, Q=64, C=10, D=4, n_min=1, n_max=20, seed=0)
    matcher_case("matcher_d4_tall", B=3, Q=16, C=10, D=4, n_min=17, n_max=30, seed=1)   # more targets than queries
    matcher_case("matcher_d7_q50", B=2, Q=50, C=10, D=7, n_min=1, n_max=12, seed=2)
    matcher_case("matcher_d5_l1", B=2, Q=20, C=5, D=5, n_min=1, n_max=8, seed=3)         # L1-only fallback
