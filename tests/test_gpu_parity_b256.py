"""GPU (-m gpu): bf16 parity on the BENCH configuration itself - B=256 frames of 3x256x256, bf16
(BASELINE.json configs[1]; every BENCH/SCALE number is quoted on it).

Reference behaviour: models/automoe.py:189-233 under torch.autocast (inference/run_automoe.py:51).  The oracle
(oracle/automoe_oracle.py, pinned to the reference's golden vectors) runs on the same device
  (a) under torch.autocast('cuda', bfloat16)  -> "the reference bf16 path"
  (b) in fp32 with TF32 off                   -> the exact answer both bf16 paths approximate
and every output key plus the full-resolution logits are compared with max|a-b| / max|b|.
The measured errors are written to gpurun_out/r2_parity_b256.json (copied to profiles/ and committed).
Routing: every frame whose top-1 expert differs from the fp32 oracle is reported with its fp32 logit gap.
"""
import json
from pathlib import Path

import pytest
import torch

from _util import build_b200_model, rel_err, rel_l2
from oracle import automoe_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2          # north_star: "within 1e-2 relative (bf16)"
KEYS = ["waypoints", "speed_seq", "expert_weights", "context_features", "combined_features", "gate_logits"]


def _outputs(o):
    d = {k: o[k].float() for k in KEYS}
    d["det_class_logits"] = o["expert_outputs"][0]["class_logits"].float()
    d["det_bbox_deltas"] = o["expert_outputs"][0]["bbox_deltas"].float()
    d["seg_logits_fullres"] = o["expert_outputs"][1]
    d["drivable_logits_fullres"] = o["expert_outputs"][2]
    return d


def test_bf16_parity_at_bench_config():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = build_b200_model(DEV, "auto")
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    sdd_cl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sdd.items()}
    B = 256
    batch = {k: v.to(DEV) for k, v in synth.synth_batch(B, 256, 256, seed=1).items()}
    batch_cl = dict(batch, image=batch["image"].contiguous(memory_format=torch.channels_last))
    cfg = synth.CONFIG_3EXPERT
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ours = _outputs(m(batch))
        # the captured graph bench.py replays must be the same computation
        g = m.capture(batch)
        replay = _outputs(g())
        torch.cuda.synchronize()
        graph_equal = all(torch.equal(ours[k], replay[k]) for k in ours)
        del g, replay
        ref32 = _outputs(O.automoe_forward(sdd, batch, cfg))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref16 = _outputs(O.automoe_forward(sdd, batch, cfg))
            # the reference's OWN bf16 path a second time, with other cuDNN algorithms (channels_last + benchmark): how far
            # two runs of the same reference arithmetic are apart once the accumulation order changes
            torch.backends.cudnn.benchmark = True
            ref16b = _outputs(O.automoe_forward(sdd_cl, batch_cl, cfg))
            torch.backends.cudnn.benchmark = False
    report = {"config": "B=256, 3x256x256, bf16 (BASELINE.json configs[1])", "tolerance": TOL,
              "metric": "max|a-b| / max|b| (BASELINE.md section 5); *_l2 = ||a-b|| / ||b||",
              "graph_replay_bit_identical_to_eager": graph_equal, "keys": {}}
    for k in ours:
        a, r16, r16b, r32 = ours[k].float(), ref16[k].float(), ref16b[k].float(), ref32[k].float()
        report["keys"][k] = {
            "ours_vs_reference_bf16": rel_err(a, r16), "ours_vs_fp32": rel_err(a, r32),
            "reference_bf16_vs_fp32": rel_err(r16, r32),
            "reference_bf16_vs_itself_other_cudnn_algos": rel_err(r16b, r16),
            "ours_vs_reference_bf16_l2": rel_l2(a, r16), "ours_vs_fp32_l2": rel_l2(a, r32),
            "reference_bf16_vs_fp32_l2": rel_l2(r16, r32),
        }
    # routing: top-1 expert against the fp32 oracle and against the reference bf16 path, flips listed with their gap
    top32 = ref32["gate_logits"].topk(2, dim=1).values
    gap = (top32[:, 0] - top32[:, 1])

    def flips(a, b):
        idx = (a["gate_logits"].argmax(1) != b["gate_logits"].argmax(1)).nonzero().flatten().tolist()
        return [{"frame": i, "fp32_logit_gap": float(gap[i])} for i in idx]
    report["routing"] = {
        "frames": B,
        "ours_vs_fp32_flips": flips(ours, ref32),
        "reference_bf16_vs_fp32_flips": flips(ref16, ref32),
        "ours_vs_reference_bf16_flips": flips(ours, ref16),
        "fp32_logit_gap_min": float(gap.min()), "fp32_logit_gap_median": float(gap.median()),
        "gate_logit_abs_err_ours_vs_fp32": float((ours["gate_logits"] - ref32["gate_logits"]).abs().max()),
        "gate_logit_abs_err_reference_bf16_vs_fp32": float((ref16["gate_logits"] - ref32["gate_logits"]).abs().max()),
    }
    report["gate"] = {
        "model_outputs (waypoints, speed_seq, expert_weights, context/combined features, gate_logits)": "max-norm <= 1e-2 vs the reference bf16 path (plain)",
        "every key, both norms": "ours_vs_fp32 <= max(1e-2, reference_bf16_vs_fp32): at least as close to the exact answer as the "
        "reference's own bf16 path",
        "expert logits (18 bf16 layers deep)": "max-norm <= max(1e-2, 1.5 x reference_bf16_vs_fp32), L2 <= max(1e-2, 1.25 x "
        "reference_bf16_vs_fp32_l2): two bf16 evaluations of this network are 1.3-1.5e-2 apart in max-norm even when they are the "
        "SAME reference arithmetic with another cuDNN algorithm (reference_bf16_vs_itself_other_cudnn_algos), so a plain 1e-2 "
        "against the reference bf16 path is not attainable by any independent implementation (DESIGN.md section 4)",
    }
    report["versions"] = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0)}
    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "r2_parity_b256.json").write_text(json.dumps(report, indent=1))
    print(json.dumps(report["keys"], indent=1))

    assert graph_equal
    for k in KEYS:      # model outputs: the stated gate, plain
        assert report["keys"][k]["ours_vs_reference_bf16"] <= TOL, (k, report["keys"][k])
        assert report["keys"][k]["ours_vs_reference_bf16_l2"] <= TOL, (k, report["keys"][k])
    for k, v in report["keys"].items():
        # every key, both norms: at least as close to the exact (fp32) answer as the reference's own bf16 path is
        assert v["ours_vs_fp32"] <= max(TOL, v["reference_bf16_vs_fp32"]), (k, v)
        assert v["ours_vs_fp32_l2"] <= max(TOL, v["reference_bf16_vs_fp32_l2"]), (k, v)
        # expert logits (18 bf16 layers deep): distance to the reference bf16 path bounded by that path's own error
        assert v["ours_vs_reference_bf16"] <= max(TOL, 1.5 * v["reference_bf16_vs_fp32"]), (k, v)
        assert v["ours_vs_reference_bf16_l2"] <= max(TOL, 1.25 * v["reference_bf16_vs_fp32_l2"]), (k, v)
    # routing: bit-exact top-1 against the fp32 oracle, or a reported flip on a frame whose fp32 logit gap is below the
    # bf16 error of the reference's own path
    noise = 2 * report["routing"]["gate_logit_abs_err_reference_bf16_vs_fp32"]
    for f in report["routing"]["ours_vs_fp32_flips"]:
        assert f["fp32_logit_gap"] <= noise, f
