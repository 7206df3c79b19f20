"""GPU (-m gpu): batched cost-matrix kernel + native LSAP against the reference's golden
assignments, the numpy oracle, and size-independent properties at BASELINE sizes."""
import numpy as np
import pytest
import torch

from oracle import matcher_oracle as MO
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["matcher_d4_q64", "matcher_d4_tall", "matcher_d7_q50", "matcher_d5_l1"]


def _dev(outputs, targets):
    return ({k: v.to(DEV) for k, v in outputs.items()}, [{k: v.to(DEV) for k, v in t.items()} for t in targets])


@pytest.mark.parametrize("name", CASES)
def test_matcher_equals_reference_golden(name, golden_dir):
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    g = np.load(golden_dir / f"{name}.npz")
    B = int(g["B"])
    outputs, targets = synth.synth_matcher_case(B, int(g["Q"]), int(g["C"]), int(g["D"]), int(g["n_min"]), int(g["n_max"]), int(g["seed"]))
    o, t = _dev(outputs, targets)
    m = HungarianMatcher()
    cost, n_tgt = m.cost_matrices(o, t)
    idx = m(o, t)
    for b in range(B):
        n = int(n_tgt[b])
        np.testing.assert_allclose(cost[b, :, :n].cpu().numpy(), g[f"cost_{b}"], rtol=0, atol=3e-6)
        assert (cost[b, :, n:] == 0).all()
        assert idx[b][0].dtype == torch.int64 and idx[b][0].device.type == "cuda"
        assert np.array_equal(idx[b][0].cpu().numpy(), g[f"rows_{b}"])     # bit-exact assignments
        assert np.array_equal(idx[b][1].cpu().numpy(), g[f"cols_{b}"])


@pytest.mark.parametrize("B,Q,C,D,nmax,seed", [(64, 64, 10, 4, 60, 21), (64, 920, 10, 4, 60, 22), (8, 196, 10, 7, 40, 23)])
def test_matcher_full_size_vs_oracle_and_properties(B, Q, C, D, nmax, seed):
    """BASELINE config 5 sizes (B=64; Q=64 for 256^2, Q=920 for 720x1280)."""
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    outputs, targets = synth.synth_matcher_case(B, Q, C, D, 1, nmax, seed)
    o, t = _dev(outputs, targets)
    m = HungarianMatcher()
    idx = m(o, t)
    ref_idx, ref_cost = MO.match(outputs["pred_logits"].numpy(), outputs["pred_boxes"].numpy(),
                                 [(x["boxes"].numpy(), x["labels"].numpy()) for x in targets])
    cost, _ = m.cost_matrices(o, t)
    cost = cost.cpu().numpy()
    for b in range(B):
        r, c = idx[b][0].cpu().numpy(), idx[b][1].cpu().numpy()
        n = targets[b]["labels"].shape[0]
        # properties: a perfect matching of min(Q, n) pairs, rows ascending, no duplicates
        assert len(r) == len(c) == min(Q, n)
        assert np.all(np.diff(r) > 0) and len(set(c.tolist())) == len(c)
        np.testing.assert_allclose(cost[b, :, :n], ref_cost[b], rtol=0, atol=5e-6)
        assert np.array_equal(r, ref_idx[b][0]) and np.array_equal(c, ref_idx[b][1])


def test_matcher_edge_cases():
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    m = HungarianMatcher()
    outputs, targets = synth.synth_matcher_case(3, 8, 5, 4, 1, 4, 31)
    targets[1] = {"boxes": torch.zeros(0, 4), "labels": torch.zeros(0, dtype=torch.int64)}  # image without objects
    o, t = _dev(outputs, targets)
    idx = m(o, t)
    assert idx[1][0].numel() == 0 and idx[1][1].numel() == 0
    # all images empty
    te = [{"boxes": torch.zeros(0, 4, device=DEV), "labels": torch.zeros(0, dtype=torch.int64, device=DEV)}] * 3
    assert all(i[0].numel() == 0 for i in m(o, te))
    # degenerate predicted boxes (w = h = 0 on top of a degenerate target) give NaN GIoU -> ValueError like scipy
    o2 = {k: v.clone() for k, v in o.items()}
    o2["pred_boxes"][0, :, 2:] = 0
    t2 = [dict(x) for x in t]
    t2[0] = {"boxes": o2["pred_boxes"][0, :2].clone(), "labels": t[0]["labels"][:1].repeat(2)}
    with pytest.raises(ValueError):
        m(o2, t2)
