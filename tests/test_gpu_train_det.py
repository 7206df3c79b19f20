"""GPU (-m gpu): detection-expert training step through the C-ABI kernels against torch
autograd, the detection training oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from _util import rel_err, rel_l2
from oracle import detection_train_oracle as DO
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 7, 9, 64), (3, 8, 6, 4), (2, 1, 1, 8), (2, 45, 80, 64), (1, 2, 3, 4)])
def test_maxpool_fwd_bwd_including_ties(shape):
    from automoe_b200.training import functional as TF
    B, H, W, C = shape
    x = torch.randn(shape, generator=_gen(1))
    x = (x * 2).round() / 2                       # many exact ties inside the 3x3 windows
    x = x.to(DEV).requires_grad_(True)
    y = TF.max_pool3x3s2(x)
    dy = torch.randn(y.shape, generator=_gen(2)).to(DEV)
    y.backward(dy)
    xr = x.detach().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    yr.backward(dy.permute(0, 3, 1, 2))
    assert torch.equal(y.detach().permute(0, 3, 1, 2), yr.detach())
    assert rel_err(x.grad.permute(0, 3, 1, 2), xr.grad) < 1e-6     # tie routing identical to torch
    # the one-pass kernel (window rescan per input pixel) and the two-pass one (arg-max taps, then comparisons) agree bit for bit
    from automoe_b200._cabi import check, ctx, lib
    from automoe_b200._ops import ptr, stream_ptr
    dev = torch.device(DEV)
    xd, dx1 = x.detach().contiguous(), torch.empty_like(x.detach())
    check(lib().amoe_maxpool3x3s2_bwd(ctx(dev), ptr(xd), ptr(dy), ptr(dx1), B, H, W, C, stream_ptr(dev)), "maxpool3x3s2_bwd")
    assert torch.equal(dx1, x.grad)


def test_add_relu_fwd_bwd():
    from automoe_b200.training import functional as TF
    a = torch.randn((3, 8, 8, 64), generator=_gen(3)).to(DEV).requires_grad_(True)
    b = torch.randn((3, 8, 8, 64), generator=_gen(4)).to(DEV).requires_grad_(True)
    dy = torch.randn((3, 8, 8, 64), generator=_gen(5)).to(DEV)
    y = TF.add_relu(a, b)
    y.backward(dy)
    ref = F.relu(a.detach() + b.detach())
    assert torch.equal(y.detach(), ref)
    g = dy * (ref > 0)
    assert torch.equal(a.grad, g) and torch.equal(b.grad, g)


@pytest.mark.parametrize("relu", [True, False])
def test_conv_bias_no_bn_fwd_bwd(relu):
    """Head convolutions: Conv2d(512,256,3,padding=1)+ReLU and Conv2d(256,14,1) (Cout not a multiple of 4)."""
    from automoe_b200.training import functional as TF
    g = _gen(6)
    for conv in (nn.Conv2d(64, 32, 3, 1, 1), nn.Conv2d(32, 14, 1)):
        conv = conv.to(DEV)
        x = torch.randn((2, conv.in_channels, 6, 10), generator=g).to(DEV)
        xn = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        y = TF.conv_bn_act(xn, conv, None, relu=relu)
        dy = torch.randn(y.shape, generator=g).to(DEV)
        y.backward(dy)
        got = (y.detach().permute(0, 3, 1, 2), xn.grad.permute(0, 3, 1, 2), conv.weight.grad.clone(), conv.bias.grad.clone())
        conv.zero_grad()
        xr = x.clone().requires_grad_(True)
        yr = conv(xr)
        yr = F.relu(yr) if relu else yr
        yr.backward(dy.permute(0, 3, 1, 2))
        for a, b in zip(got, (yr.detach(), xr.grad, conv.weight.grad, conv.bias.grad)):
            assert rel_err(a, b) < TOL, rel_err(a, b)


def _det_loss_case(B, Q, n_max, seed):
    g = _gen(seed)
    head = torch.randn((B, Q, 1, 14), generator=g)
    tcls = torch.full((B * Q,), 10, dtype=torch.int64)
    tbox = torch.zeros((B * Q, 4))
    for b in range(B):
        n = int(torch.randint(0, n_max + 1, (1,), generator=g))
        rows = torch.randperm(Q, generator=g)[:n] + b * Q
        tcls[rows] = torch.randint(0, 10, (n,), generator=g)
        tbox[rows] = torch.rand((n, 4), generator=g) * 3 - 1     # both SmoothL1 branches
    return head, tcls, tbox


@pytest.mark.parametrize("B,Q,n_max", [(4, 20, 5), (64, 920, 60), (2, 6, 6)])
def test_det_loss_fwd_bwd(B, Q, n_max):
    from automoe_b200.training.functional import _DetLoss
    head, tcls, tbox = _det_loss_case(B, Q, n_max, 7)
    h = head.to(DEV).requires_grad_(True)
    losses = _DetLoss.apply(h, tcls.to(DEV), tbox.to(DEV), 10, 2.0)
    losses[0].backward()
    hr = head.clone().requires_grad_(True)
    flat = hr.reshape(B * Q, 14)
    ce = F.cross_entropy(flat[:, :10], tcls, ignore_index=10)
    m = tcls != 10
    bx = F.smooth_l1_loss(flat[:, 10:][m], tbox[m], reduction="mean")
    (ce + 2.0 * bx).backward()
    assert abs(losses[1].item() - ce.item()) < 1e-5 * max(1, abs(ce.item())) and abs(losses[2].item() - bx.item()) < 1e-5
    assert int(losses[3].item()) == int(m.sum())
    assert rel_err(h.grad.cpu(), hr.grad) < TOL


def _b200_det_expert():
    from automoe_b200.models.automoe import create_automoe_model
    from automoe_b200.models.experts import BDDDetectionExpert
    full = synth.synth_state_dict(create_automoe_model(synth.CONFIG_3EXPERT, "cpu").state_dict(), 0)
    sd = {k[len("experts.0."):]: v for k, v in full.items() if k.startswith("experts.0.")}
    m = BDDDetectionExpert(num_classes=10, pretrained_backbone=False)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV), sd


def test_detection_training_step_matches_reference_golden(golden_dir):
    """Expert forward (train-mode BatchNorm) -> matcher -> scatter -> CE + 2*SmoothL1 -> backward: loss and the
    gradients of all 12,360,014 parameters against the unmodified reference."""
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    from automoe_b200.training.train_bdd100k import train_detection_batch
    g = np.load(golden_dir / "det_train_b3_128x160.npz")
    m, _ = _b200_det_expert()
    m.train()
    batch = DO.synth_detection_batch(int(g["B"]), int(g["H"]), int(g["W"]), int(g["n_max"]), int(g["seed"]))
    loss = train_detection_batch(m, batch, HungarianMatcher())
    loss.backward()
    assert abs(loss.item() - float(g["total_loss"])) < 1e-4 * float(g["total_loss"]), (loss.item(), float(g["total_loss"]))
    params = dict(m.named_parameters())
    assert sum(p.numel() for p in params.values()) == 12360014
    worst = 0.0
    for n, norm in zip([str(x) for x in g["grad_names"]], g["grad_norms"]):
        got = params[n].grad.double().norm().item()
        if norm < 1e-5:
            assert got < 1e-3, (n, got)
            continue
        worst = max(worst, abs(got - norm) / norm)
        assert abs(got - norm) < 2e-3 * norm, (n, got, norm)
    for key in g.files:
        if key.startswith("full__") and float(np.abs(g[key]).max()) > 1e-6:
            assert rel_err(params[key[6:]].grad.cpu(), g[key]) < 2e-3, (key, rel_err(params[key[6:]].grad.cpu(), g[key]))
    bn = m.backbone[1]
    assert rel_err(bn.running_mean.cpu(), g["bn1_running_mean"]) < TOL and rel_err(bn.running_var.cpu(), g["bn1_running_var"]) < TOL
    print("worst relative grad-norm error vs reference:", worst)


def test_detection_training_step_matches_oracle_and_assignment_bit_exact():
    """Larger ragged case vs the oracle: identical Hungarian assignment (bit-exact indices), loss within 1e-4."""
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    from automoe_b200.training.train_bdd100k import detection_losses
    m, sd = _b200_det_expert()
    m.train()
    batch = DO.synth_detection_batch(4, 192, 256, 12, 5)
    out = m(batch["image"].to(DEV))
    res = detection_losses(out, batch["bboxes"], batch["labels"], HungarianMatcher(), 10, 2.0)
    res["total_loss"].backward()
    sdd = {k: v.to(DEV).clone() for k, v in sd.items()}
    for k, v in sdd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    oref = DO.detection_forward_train(batch["image"].to(DEV), sdd)
    lref = DO.detection_loss(oref, batch["bboxes"].to(DEV), batch["labels"].to(DEV))
    lref["total_loss"].backward()
    for (r, c), (rr, rc) in zip(res["indices"], lref["indices"]):
        assert r.cpu().tolist() == list(rr) and c.cpu().tolist() == list(rc)
    assert abs(res["total_loss"].item() - lref["total_loss"].item()) < 1e-4 * abs(lref["total_loss"].item())
    assert rel_err(out["class_logits"].detach(), oref["class_logits"].detach()) < TOL
    # two correct fp32 implementations of a ReLU network disagree on the few units whose pre-activation is
    # within rounding of zero; the detection loss back-propagates from a dozen matched queries only, so one flip on
    # their path shifts early-layer gradients by ~1e-2 (measured 6e-3 on layer1 here; the golden test above holds
    # every tensor to 2e-3 against the real reference on its case) - compare in L2 with that much room
    for k in ("head.2.weight", "head.0.weight", "backbone.7.1.conv2.weight", "backbone.4.0.conv1.weight", "backbone.0.weight"):
        ours, ref = dict(m.named_parameters())[k].grad, sdd[k].grad
        assert rel_l2(ours, ref) < 2e-2 and rel_err(ours, ref) < 5e-2, (k, rel_l2(ours, ref), rel_err(ours, ref))


@pytest.mark.parametrize("shape", [(2, 4, 5, 19, 128, 160), (1, 8, 8, 3, 256, 256), (2, 3, 3, 4, 50, 70)])
def test_upsample_bilinear_fwd_bwd(shape):
    """x32 (and non-integer scale) bilinear up-sampling NHWC -> NCHW against F.interpolate, forward and adjoint."""
    from automoe_b200.training import functional as TF
    B, h, w, C, H, W = shape
    low = torch.randn((B, h, w, C), generator=_gen(8)).to(DEV).requires_grad_(True)
    dy = torch.randn((B, C, H, W), generator=_gen(9)).to(DEV)
    y = TF.upsample_bilinear_nchw(low, H, W)
    y.backward(dy)
    lr = low.detach().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yr = F.interpolate(lr, size=(H, W), mode="bilinear", align_corners=False)
    yr.backward(dy)
    assert rel_err(y.detach(), yr.detach()) < 1e-5
    assert rel_err(low.grad.permute(0, 3, 1, 2), lr.grad) < 1e-5, rel_err(low.grad.permute(0, 3, 1, 2), lr.grad)


def test_segmentation_expert_training_step_matches_oracle():
    """BDDTrainer._train_segmentation_batch (train_bdd100k_ddp.py:188-194): expert forward in train mode +
    nn.CrossEntropyLoss(ignore_index=255) (the reference's own torch loss on our differentiable output)."""
    from automoe_b200.models.automoe import create_automoe_model
    from automoe_b200.models.experts import BDDSegmentationExpert
    from oracle import automoe_oracle as O
    full = synth.synth_state_dict(create_automoe_model(synth.CONFIG_3EXPERT, "cpu").state_dict(), 0)
    sd = {k[len("experts.1."):]: v for k, v in full.items() if k.startswith("experts.1.")}
    m = BDDSegmentationExpert(num_classes=19, pretrained_backbone=False)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).train()
    g = _gen(13)
    x = torch.randn((2, 3, 64, 96), generator=g).to(DEV)
    mask = torch.randint(0, 19, (2, 64, 96), generator=g)
    mask[torch.rand((2, 64, 96), generator=g) < 0.1] = 255
    mask = mask.to(DEV)
    out = m(x)
    assert out.shape == (2, 19, 64, 96) and out.requires_grad
    loss = F.cross_entropy(out, mask, ignore_index=255)
    loss.backward()
    sdd = {k: v.to(DEV).clone() for k, v in sd.items()}
    for k, v in sdd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    low = O.expert_head(DO.trunk_train(x, sdd, "backbone"), sdd, "decoder")
    ref = F.interpolate(low, size=x.shape[-2:], mode="bilinear", align_corners=False)
    lref = F.cross_entropy(ref, mask, ignore_index=255)
    lref.backward()
    assert rel_err(out.detach(), ref.detach()) < TOL
    assert abs(loss.item() - lref.item()) < 1e-4 * abs(lref.item())
    params = dict(m.named_parameters())
    for k in ("decoder.2.weight", "decoder.2.bias", "decoder.0.weight", "backbone.7.1.conv2.weight", "backbone.0.weight"):
        assert rel_l2(params[k].grad, sdd[k].grad) < 5e-3, (k, rel_l2(params[k].grad, sdd[k].grad))
