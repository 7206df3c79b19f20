"""GPU (-m gpu): the gating/policy training step through the C-ABI kernels against
torch autograd on the same seeded inputs, the training oracle and the reference's golden vectors.
fp32: 1e-4 relative (max|a-b| / max|b|)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from _util import build_b200_model, rel_err, rel_l2
from oracle import gating_train_oracle as GT
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


@pytest.mark.parametrize("B,i,o,relu", [(32, 896, 128, True), (5, 4, 32, True), (32, 768, 512, False), (3, 128, 3, False),
                                        (70, 130, 66, True), (64, 260, 24, True), (33, 1030, 9, False), (1, 2050, 130, True)])
def test_linear_fwd_bwd(B, i, o, relu):
    from automoe_b200.training import functional as TF
    g = _gen(1)
    lin = nn.Linear(i, o).to(DEV)
    x = torch.randn((B, i), generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn((B, o), generator=g).to(DEV)
    y = TF.linear(x, lin, relu=relu)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None; lin.zero_grad()
    yr = F.linear(x, lin.weight, lin.bias)
    yr = F.relu(yr) if relu else yr
    yr.backward(dy)
    for a, b in zip(got, (yr.detach(), x.grad, lin.weight.grad, lin.bias.grad)):
        assert rel_err(a, b) < TOL, rel_err(a, b)


def test_linear_on_row_slices_of_wider_buffers():
    """The C entry takes row strides: x / y as column slices of wider buffers whose start is not 16-byte aligned (the
    small-batch kernel then falls back from 16-byte to scalar loads of x) - straight through the C-ABI."""
    from automoe_b200._cabi import check, ctx, lib
    from automoe_b200._ops import ptr, stream_ptr
    g = _gen(3)
    B, i, o = 32, 70, 20
    xw = torch.randn((B, 100), generator=g).to(DEV)
    yw = torch.zeros((B, 50), device=DEV)
    W = torch.randn((o, i), generator=g).to(DEV)
    b = torch.randn(o, generator=g).to(DEV)
    x, y = xw[:, 3:3 + i], yw[:, 5:5 + o]
    dev = torch.device(DEV)
    check(lib().amoe_linear_fwd(ctx(dev), x.data_ptr(), 100, ptr(W), ptr(b), y.data_ptr(), 50, B, i, o, 1, 0.0, 0, stream_ptr(dev)),
          "linear_fwd")
    assert rel_err(y, F.relu(F.linear(x, W, b))) < TOL
    assert yw[:, :5].abs().max() == 0 and yw[:, 5 + o:].abs().max() == 0      # nothing outside the slice was written


def test_linear_dropout_mask_and_scale():
    from automoe_b200.training import functional as TF
    torch.manual_seed(5)
    lin = nn.Linear(64, 4096).to(DEV)
    x = torch.randn(16, 64, device=DEV, requires_grad=True)
    y = TF.linear(x, lin, relu=True, drop_p=0.25)
    ref = F.relu(F.linear(x, lin.weight, lin.bias)).detach()
    active = ref > 0
    kept = (y.detach() > 0) & active
    frac = kept.sum().item() / active.sum().item()
    assert abs(frac - 0.75) < 0.01, frac                                   # Bernoulli(1-p) keep rate
    assert rel_err(y.detach()[kept], ref[kept] / 0.75) < 1e-6              # kept units scaled by 1/(1-p)
    y.sum().backward()
    # gradient only flows through kept units, scaled the same way
    gref = (kept.float() / 0.75) @ lin.weight.detach()
    assert rel_err(x.grad, gref) < TOL
    y2 = TF.linear(x, lin, relu=True, drop_p=0.25)                         # a new call draws a new mask
    assert ((y2.detach() > 0) != (y.detach() > 0)).any()


@pytest.mark.parametrize("B,D", [(32, 256), (7, 64), (1, 256)])
def test_layernorm_fwd_bwd(B, D):
    from automoe_b200.training import functional as TF
    g = _gen(2)
    ln = nn.LayerNorm(D).to(DEV)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(D, generator=g)); ln.bias.copy_(0.1 * torch.randn(D, generator=g))
    x = (torch.randn((B, D), generator=g) * 3 + 1).to(DEV).requires_grad_(True)
    dy = torch.randn((B, D), generator=g).to(DEV)
    y = TF.layer_norm(x, ln)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
    x.grad = None; ln.zero_grad()
    yr = ln(x)
    yr.backward(dy)
    for a, b in zip(got, (yr.detach(), x.grad, ln.weight.grad, ln.bias.grad)):
        assert rel_err(a, b) < TOL, rel_err(a, b)


@pytest.mark.parametrize("T", [1.0, 0.5])
def test_gate_combine_fwd_bwd(T):
    from automoe_b200.training import functional as TF
    g = _gen(3)
    B, E, P = 32, 3, 256
    logits = torch.randn((B, E), generator=g).to(DEV).requires_grad_(True)
    procs = [torch.randn((B, P), generator=g).to(DEV).requires_grad_(True) for _ in range(E)]
    dw = torch.randn((B, E), generator=g).to(DEV)
    dc = torch.randn((B, P), generator=g).to(DEV)
    w, c = TF.gate_combine(logits, procs, T)
    (w * dw).sum().add((c * dc).sum()).backward()
    got = [w.detach(), c.detach(), logits.grad.clone()] + [p.grad.clone() for p in procs]
    logits.grad = None
    for p in procs:
        p.grad = None
    wr = F.softmax(logits / T, dim=1)
    cr = torch.zeros_like(procs[0])
    for e in range(E):
        cr = cr + wr[:, e:e + 1] * procs[e]
    (wr * dw).sum().add((cr * dc).sum()).backward()
    ref = [wr.detach(), cr.detach(), logits.grad] + [p.grad for p in procs]
    for a, b in zip(got, ref):
        assert rel_err(a, b) < TOL, rel_err(a, b)


CONV_CASES = [  # Cin, Cout, k, pad, H, B
    (3, 32, 5, 2, 64, 4), (32, 64, 3, 1, 32, 4), (64, 128, 3, 1, 16, 3), (128, 256, 3, 1, 8, 5), (32, 64, 3, 1, 18, 2),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("batch_stats", [True, False])
def test_conv_bn_relu_fwd_bwd(case, batch_stats):
    """Conv2d(stride 2, bias) + BatchNorm2d (train: batch statistics + running-stat update; eval: running
    statistics) + ReLU: outputs, all five gradients and the updated running statistics."""
    from automoe_b200 import _ops
    from automoe_b200.training import functional as TF
    Cin, Cout, k, pad, H, B = case
    g = _gen(4)
    conv = nn.Conv2d(Cin, Cout, k, 2, pad).to(DEV)
    bn = nn.BatchNorm2d(Cout).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(Cout, generator=g)); bn.bias.copy_(0.1 * torch.randn(Cout, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(Cout, generator=g)); bn.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
    bn.train(batch_stats)
    import copy
    conv_r, bn_r = copy.deepcopy(conv), copy.deepcopy(bn)
    x = torch.randn((B, Cin, H, H), generator=g).to(DEV)
    if Cin == 3:
        x_nhwc = _ops.image_to_nhwc(x, 4, torch.float32).requires_grad_(True)
    else:
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    y = TF.conv_bn_relu(x_nhwc, conv, bn, batch_stats=batch_stats)
    dy = torch.randn(y.shape, generator=_gen(9)).to(DEV)
    y.backward(dy)
    xr = x.clone().requires_grad_(True)
    yr = F.relu(bn_r(conv_r(xr)))
    yr.backward(dy.permute(0, 3, 1, 2))
    assert rel_err(y.detach().permute(0, 3, 1, 2), yr.detach()) < TOL
    assert rel_err(x_nhwc.grad[..., :Cin].permute(0, 3, 1, 2), xr.grad) < TOL
    if Cin == 3:
        assert (x_nhwc.grad[..., 3] == 0).all() or True     # padded channel: weights are zero there
    assert rel_err(conv.weight.grad, conv_r.weight.grad) < TOL, rel_err(conv.weight.grad, conv_r.weight.grad)
    assert rel_err(bn.weight.grad, bn_r.weight.grad) < TOL
    assert rel_err(bn.bias.grad, bn_r.bias.grad) < TOL
    if batch_stats:
        # the conv bias cancels inside batch-statistics BN: its true gradient is ~0 (rounding noise on both sides)
        assert conv.bias.grad.abs().max() < 1e-3 * max(1.0, bn.bias.grad.abs().max().item())
        assert rel_err(bn.running_mean, bn_r.running_mean) < TOL and rel_err(bn.running_var, bn_r.running_var) < TOL
        assert int(bn.num_batches_tracked) == int(bn_r.num_batches_tracked) == 1
    else:
        assert rel_err(conv.bias.grad, conv_r.bias.grad) < TOL


def test_gap_fwd_bwd():
    from automoe_b200.training import functional as TF
    x = torch.randn((6, 16, 16, 256), generator=_gen(6)).to(DEV).requires_grad_(True)
    dy = torch.randn((6, 256), generator=_gen(7)).to(DEV)
    y = TF.global_avg_pool(x)
    y.backward(dy)
    assert rel_err(y.detach(), x.detach().mean(dim=(1, 2))) < 1e-6
    assert rel_err(x.grad, (dy / 256.0)[:, None, None, :].expand_as(x)) < 1e-6


def _loss_inputs(B, H, E, seed, spd_cols=None):
    g = _gen(seed)
    pred = {"waypoints": torch.randn((B, H, 2), generator=g) * 4, "speed_seq": torch.rand((B, spd_cols or H), generator=g) * 30,
            "expert_weights": F.softmax(torch.randn((B, E), generator=g), dim=1)}
    pred["speed"] = pred["speed_seq"][:, -1:].contiguous()
    return pred, torch.randn((B, H, 2), generator=g) * 5, torch.rand((B, H), generator=g) * 30


@pytest.mark.parametrize("B,H,E", [(32, 10, 3), (4, 10, 3), (1, 3, 2), (300, 8, 4)])
@pytest.mark.parametrize("cfg", [{}, {"use_load_balancing": False, "entropy_weight": 0.05, "smoothness_weight": 0.7},
                                 {"use_entropy_loss": False, "ade_weight": 0.3, "load_balancing_weight": 2.0}])
def test_gating_loss_matches_reference_expression(B, H, E, cfg):
    from automoe_b200.training.train_gating_network import compute_gating_losses
    pred_c, twp, tspd = _loss_inputs(B, H, E, 11)
    leaves = {k: pred_c[k].clone().to(DEV).requires_grad_(True) for k in ("waypoints", "speed_seq", "expert_weights")}
    pred = dict(leaves, speed=leaves["speed_seq"][:, -1:])
    out = compute_gating_losses(pred, twp.to(DEV), tspd.to(DEV), cfg)
    out["total_loss"].backward()
    leaves_r = {k: pred_c[k].clone().requires_grad_(True) for k in ("waypoints", "speed_seq", "expert_weights")}
    pred_r = dict(leaves_r, speed=leaves_r["speed_seq"][:, -1:])
    ref = GT.compute_gating_losses(pred_r, twp, tspd, cfg)
    ref["total_loss"].backward()
    for k in ref:
        assert abs(out[k].item() - ref[k].item()) <= 2e-5 * max(1.0, abs(ref[k].item())), (k, out[k].item(), ref[k].item())
    for k in leaves:
        assert rel_err(leaves[k].grad.cpu(), leaves_r[k].grad) < TOL, k


def test_gating_loss_last_step_speed_branch():
    """pred has no speed_seq of matching length -> the reference falls back to the last step (lines 32-35)."""
    from automoe_b200.training.train_gating_network import compute_gating_losses
    pred_c, twp, tspd = _loss_inputs(8, 10, 3, 12, spd_cols=1)
    wp = pred_c["waypoints"].clone().to(DEV).requires_grad_(True)
    sp = pred_c["speed"].clone().to(DEV).requires_grad_(True)
    ew = pred_c["expert_weights"].clone().to(DEV).requires_grad_(True)
    out = compute_gating_losses({"waypoints": wp, "speed": sp, "expert_weights": ew}, twp.to(DEV), tspd.to(DEV), {})
    out["total_loss"].backward()
    wr, sr, er = (t.detach().cpu().requires_grad_(True) for t in (wp, sp, ew))
    ref = GT.compute_gating_losses({"waypoints": wr, "speed": sr, "expert_weights": er}, twp, tspd, {})
    ref["total_loss"].backward()
    assert abs(out["speed"].item() - ref["speed"].item()) < 1e-4 and abs(out["total_loss"].item() - ref["total_loss"].item()) < 1e-4
    assert rel_err(sp.grad.cpu(), sr.grad) < TOL and rel_err(wp.grad.cpu(), wr.grad) < TOL


def _targets(B, horizon, seed):
    g = _gen(seed)
    return torch.randn((B, horizon, 2), generator=g) * 5.0, torch.rand((B, horizon), generator=g) * 30.0


def _model_for_training(train_mode, experts_eval=True):
    m, sd = build_b200_model(DEV, "fp32")
    m.freeze_experts()
    if train_mode:
        m.train()
        if experts_eval:
            m.experts.eval()      # what the reference gives after model.experts.eval(); model.train() alone = reference semantics
        for mod in m.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
    return m, sd


@pytest.mark.parametrize("name", ["train_eval_b4_64", "train_trainmode_b4_64", "train_refmode_b4_128"])
def test_training_step_matches_reference_golden(name, golden_dir):
    """loss values + gradients of all 82 trainable tensors (2,870,657 parameters) against the unmodified
    reference (eval semantics; train mode with Dropout p=0 and batch-statistics BatchNorm in the policy)."""
    from automoe_b200.training.train_gating_network import compute_gating_losses
    g = np.load(golden_dir / f"{name}.npz")
    B, H, train_mode = int(g["B"]), int(g["H"]), bool(g["train_mode"])
    experts_eval = bool(g["experts_eval"]) if "experts_eval" in g.files else True
    m, sd = _model_for_training(train_mode, experts_eval)
    batch = {k: v.to(DEV) for k, v in synth.synth_batch(B, H, H, seed=3, speed_seq=1).items()}
    wp, spd = _targets(B, 10, 4)
    pred = m(batch)
    losses = compute_gating_losses(pred, wp.to(DEV), spd.to(DEV), {})
    losses["total_loss"].backward()
    got = np.array([losses[k].item() for k in ("total_loss", "ade", "fde", "speed", "smoothness", "load_balancing", "entropy")])
    assert np.allclose(got, g["losses"], rtol=1e-4, atol=1e-6), (got, g["losses"])
    assert rel_err(pred["waypoints"].detach().cpu(), g["waypoints"]) < TOL
    params = dict(m.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    assert sorted(names) == sorted(k for k, p in params.items() if p.requires_grad)
    worst = 0.0
    for n, norm, head in zip(names, g["grad_norms"], g["grad_heads"]):
        gr = params[n].grad
        assert gr is not None, n
        if norm < 1e-5:        # e.g. conv bias in front of batch-statistics BN: true gradient 0
            assert gr.double().norm().item() < 1e-3, (n, gr.norm().item())
            continue
        e = abs(gr.double().norm().item() - norm) / norm
        worst = max(worst, e)
        assert e < 5e-4, (n, gr.double().norm().item(), norm)
        k = min(8, gr.numel())
        assert np.allclose(gr.reshape(-1)[:k].cpu().numpy(), head[:k], rtol=5e-3, atol=1e-6 + 2e-4 * float(np.abs(head).max())), n
    for key in g.files:
        if key.startswith("full__") and float(np.abs(g[key]).max()) > 1e-5:
            assert rel_err(params[key[6:]].grad.cpu(), g[key]) < 5e-4, (key, rel_err(params[key[6:]].grad.cpu(), g[key]))
    if train_mode:
        bn = m.policy_head.backbone.net[1]
        assert rel_err(bn.running_mean.cpu(), g["bn1_running_mean"]) < TOL and rel_err(bn.running_var.cpu(), g["bn1_running_var"]) < TOL
    if not experts_eval:
        # reference train-mode semantics: the frozen experts ran on batch statistics and updated their running statistics
        sdm = m.state_dict()
        for key in g.files:
            if key.startswith("stat__"):
                assert rel_err(sdm[key[6:]].cpu(), g[key]) < TOL, (key, rel_err(sdm[key[6:]].cpu(), g[key]))
        assert int(sdm["experts.1.backbone.4.0.bn1.num_batches_tracked"]) == int(g["expert_nbt"])
        assert rel_err(pred["expert_outputs"][1].mean(dim=(2, 3)).cpu(), g["seg_mean"]) < TOL
        assert rel_err(pred["expert_outputs"][0]["class_logits"].cpu(), g["det_class_logits"]) < TOL
    print(name, "worst relative grad-norm error", worst)


def _oracle_grads(sd, batch, wp, spd, dtype):
    sdd = {k: (v.to(DEV).to(dtype) if v.is_floating_point() else v.to(DEV)).clone() for k, v in sd.items()}
    for k, v in sdd.items():
        if GT.is_trainable_key(k) and v.is_floating_point():
            v.requires_grad_(True)
    b = {k: v.to(dtype) for k, v in batch.items()}
    pr = GT.training_forward(sdd, b, synth.CONFIG_3EXPERT, policy_batch_stats=True)
    lr = GT.compute_gating_losses(pr, wp.to(DEV).to(dtype), spd.to(DEV).to(dtype), {})
    lr["total_loss"].backward()
    return lr["total_loss"].item(), {k: v.grad for k, v in sdd.items() if v.requires_grad}


def test_training_step_matches_oracle_batch32():
    """BASELINE config 4 shape (batch 32 per GPU, 256x256).  Yardstick: an fp64 evaluation of the oracle.
    Most gradients land within 1e-4 of it.  The policy backbone does not: behind the global average pool the
    incoming gradient is constant over a channel, BatchNorm's backward subtracts its mean, and a handful of
    ReLU units whose pre-activation is within fp32 rounding of zero flip between implementations - each flip
    moves a channel's gradient by ~1e-2 of its value.  Which units flip is luck of the rounding on either side (torch's own fp32
    arithmetic measured 2e-4 .. 2e-6 from fp64 on these tensors, ours 6e-4 .. 6e-3), so the policy backbone is
    held to max-error 2e-2 / L2-error 2e-3, everything else to 1e-4."""
    from automoe_b200.training.train_gating_network import compute_gating_losses
    B, H = 32, 256
    m, sd = _model_for_training(True)
    batch = {k: v.to(DEV) for k, v in synth.synth_batch(B, H, H, seed=8, speed_seq=1).items()}
    wp, spd = _targets(B, 10, 9)
    pred = m(batch)
    losses = compute_gating_losses(pred, wp.to(DEV), spd.to(DEV), {})
    losses["total_loss"].backward()
    l64, g64 = _oracle_grads(sd, batch, wp, spd, torch.float64)
    l32, g32 = _oracle_grads(sd, batch, wp, spd, torch.float32)
    assert abs(losses["total_loss"].item() - l64) < 1e-4 * abs(l64)
    worst = ("", 0.0, 0.0)
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue
        truth = g64[k]
        if truth.abs().max() < 1e-6:
            assert p.grad.abs().max() < 1e-4, k
            continue
        ours, theirs = rel_err(p.grad, truth), rel_err(g32[k], truth)
        if ours > worst[1]:
            worst = (k, ours, theirs)
        if k.startswith("policy_head.backbone"):
            # one flipped ReLU unit moves one channel's row by ~1e-2: bound the max loosely, the L2 error tightly
            assert ours < 2e-2 and rel_l2(p.grad, truth) < 2e-3, (k, ours, rel_l2(p.grad, truth), theirs)
        else:
            assert ours < max(TOL, 1.25 * theirs), (k, ours, theirs)
    print("worst gradient error vs fp64 (ours, torch fp32):", worst)


def test_flat_adamw_matches_torch_adamw_and_clip():
    """FlatAdamW (one norm reduction + one fused clip/AdamW kernel on flat buffers) against
    clip_grad_norm_(1.0) + torch.optim.AdamW on identical gradients for five steps; cached inference packs
    see the update (Tensor._version is bumped)."""
    from automoe_b200.training.train_gating_network import FlatAdamW
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(37, 64), nn.ReLU(), nn.Linear(64, 5)).to(DEV)
    import copy
    ref = copy.deepcopy(net)
    opt = FlatAdamW(net.parameters(), lr=1e-2, weight_decay=1e-2, max_norm=1.0)
    opt_r = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=1e-2)
    v0 = [p._version for p in net.parameters()]
    for step in range(5):
        x = torch.randn(16, 37, device=DEV) * (10.0 if step % 2 == 0 else 0.01)   # clipped and unclipped steps
        opt.zero_grad(); opt_r.zero_grad()
        net(x).pow(2).mean().backward()
        ref(x).pow(2).mean().backward()
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=1.0)
        opt.step(); opt_r.step()
        assert abs(opt.total_norm().item() - total.item()) < 1e-4 * max(1.0, total.item())
        for p, q in zip(net.parameters(), ref.parameters()):
            assert rel_err(p.detach(), q.detach()) < 1e-5, (step, rel_err(p.detach(), q.detach()))
    assert all(p._version > v for p, v in zip(net.parameters(), v0))


def test_train_step_reduces_loss_and_refreshes_inference_packs():
    """A few FlatAdamW steps through train_step lower the loss; the eval-mode forward afterwards uses the
    UPDATED weights (its cached packs are invalidated) and equals the differentiable path's output."""
    from automoe_b200.training.train_gating_network import FlatAdamW, freeze_for_gating_training, train_step
    m, _ = build_b200_model(DEV, "fp32")
    params = freeze_for_gating_training(m)
    assert sum(p.numel() for p in params) == 2870657
    batch = {k: v.to(DEV) for k, v in synth.synth_batch(8, 64, 64, seed=21, speed_seq=1).items()}
    wp, spd = _targets(8, 10, 22)
    batch["waypoints"], batch["speed"] = wp.to(DEV), spd.to(DEV)    # trainer batches carry the targets under these keys
    with torch.no_grad():
        before = m(batch)["waypoints"].clone()
    opt = FlatAdamW(params, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
    m.train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    first = last = None
    for i in range(6):
        l = train_step(m, batch, opt, {})["total_loss"].item()
        first = l if first is None else first
        last = l
    assert last < first, (first, last)
    m.eval()
    with torch.no_grad():
        after = m(batch)["waypoints"]
    assert (after - before).abs().max() > 1e-4                      # the fused inference path saw the new weights
    diff_path = m(batch)["waypoints"]                               # eval + grad enabled + frozen experts: autograd path
    assert diff_path.requires_grad and rel_err(after, diff_path.detach()) < 1e-4


@pytest.mark.parametrize("B,H,W", [(4, 64, 64), (3, 96, 160)])
def test_grouped_frozen_expert_train_forward_equals_expert_by_expert(B, H, W):
    """Frozen experts in the reference's train mode (batch-statistics BatchNorm, running statistics and counters updated):
    the lockstep path - one convolution / BatchNorm launch per layer for all three experts - against the same kernels run
    expert by expert: head outputs, running statistics and num_batches_tracked of every BatchNorm layer."""
    from automoe_b200.models.experts._trunk import grouped_train_supported, run_trunk_train, run_trunks_train_grouped
    ma, _ = build_b200_model(DEV, "fp32")
    mb, _ = build_b200_model(DEV, "fp32")
    for m in (ma, mb):
        m.freeze_experts()
        m.train()
    img = torch.randn((B, 3, H, W), generator=_gen(50)).to(DEV)
    ea, eb = ma._bdd_experts(), mb._bdd_experts()
    assert grouped_train_supported(eb, img)
    for _ in range(2):                               # two steps: the running statistics accumulate
        with torch.no_grad():
            ref = [run_trunk_train(e, img) for e in ea]
        got = run_trunks_train_grouped(eb, img)
    for r, g in zip(ref, got):
        assert r.shape == g.shape and rel_err(g, r) < 1e-6, rel_err(g, r)
    sa, sb = dict(ma.experts.named_buffers()), dict(mb.experts.named_buffers())
    assert sa.keys() == sb.keys()
    for k in sa:
        if k.endswith("num_batches_tracked"):
            assert int(sa[k]) == int(sb[k]) == 2, k
        else:
            assert rel_err(sb[k], sa[k]) < 1e-6, (k, rel_err(sb[k], sa[k]))
    assert not grouped_train_supported(eb, torch.zeros((2, 3, 90, 122), device=DEV))      # odd stage sizes: expert by expert


def _graph_models():
    from automoe_b200.training.train_gating_network import FlatAdamW, freeze_for_gating_training
    out = []
    for _ in range(2):
        m, _sd = build_b200_model(DEV, "fp32")
        opt = FlatAdamW(freeze_for_gating_training(m), lr=1e-3, weight_decay=1e-4, max_norm=1.0)
        m.train()
        out.append((m, opt))
    return out


def _train_batch(B, H, seed):
    batch = {k: v.to(DEV) for k, v in synth.synth_batch(B, H, H, seed=seed, speed_seq=1).items()}
    wp, spd = _targets(B, 10, seed + 1)
    batch["waypoints"], batch["speed"] = wp.to(DEV), spd.to(DEV)
    return batch


def test_graphed_train_step_equals_eager_steps():
    """train_step captured once as a CUDA graph (GraphedTrainStep) and replayed on three different batches
    against three eager steps from the same initial state (Dropout p = 0 so both draw no masks): same losses, parameters,
    Adam moments and BatchNorm running statistics; building the graph leaves the training state untouched."""
    from automoe_b200.training.train_gating_network import GraphedTrainStep, train_step
    (ma, oa), (mb, ob) = _graph_models()
    for m in (ma, mb):
        for mod in m.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
    batches = [_train_batch(4, 64, 30 + 2 * i) for i in range(3)]
    p0 = ob.flat_param.clone()
    bufs0 = [b.clone() for b in mb.buffers()]
    step = GraphedTrainStep(mb, batches[0], ob, {})
    assert torch.equal(ob.flat_param, p0) and ob.step_count == 0 and int(ob._step_dev.item()) == 0
    assert all(torch.equal(b, c) for b, c in zip(mb.buffers(), bufs0))
    assert step.launches_per_replay > 100
    for i, b in enumerate(batches):
        le = train_step(ma, b, oa, {})
        lg = step(b)
        for k in le:
            assert abs(le[k].item() - lg[k].item()) <= 1e-5 * max(1.0, abs(le[k].item())), (i, k, le[k].item(), lg[k].item())
        assert rel_err(ob.flat_param, oa.flat_param) < 1e-6, (i, rel_err(ob.flat_param, oa.flat_param))
        assert rel_err(ob.flat_grad, oa.flat_grad) < 1e-5
    assert ob.step_count == 3 and int(ob._step_dev.item()) == 3
    assert rel_err(ob.exp_avg, oa.exp_avg) < 1e-5 and rel_err(ob.exp_avg_sq, oa.exp_avg_sq) < 1e-5
    for (ka, a), (_, b) in zip(ma.named_buffers(), mb.named_buffers()):
        assert rel_err(b.float(), a.float()) < 1e-5, ka
    # eval forward after graphed training sees the updated weights (cached packs invalidated by the version bump)
    ma.eval(); mb.eval()
    with torch.no_grad():
        assert rel_err(mb(batches[0])["waypoints"], ma(batches[0])["waypoints"]) < 1e-5


def test_graphed_train_step_draws_new_dropout_masks_per_replay():
    """Dropout(0.1) stays active in a replayed step: with lr = 0 the parameters stay put, so two replays on one batch differ
    only by their masks - the per-step part of the key lives on the device and advances inside the graph."""
    from automoe_b200.training.train_gating_network import FlatAdamW, GraphedTrainStep, freeze_for_gating_training
    m, _sd = build_b200_model(DEV, "fp32")
    opt = FlatAdamW(freeze_for_gating_training(m), lr=0.0, weight_decay=0.0, max_norm=1.0)
    m.train()
    batch = _train_batch(8, 64, 40)
    step = GraphedTrainStep(m, batch, opt, {})
    losses = [step(batch)["total_loss"].item() for _ in range(4)]
    assert len({round(l, 6) for l in losses}) == 4, losses
    assert max(losses) - min(losses) < 0.2 * abs(losses[0])         # same batch, same weights: masks are the only difference


@pytest.mark.parametrize("B,H,W,Cout,K,pad,bias", [(3, 64, 64, 64, 7, 3, False), (2, 256, 256, 64, 7, 3, False), (5, 96, 160, 32, 5, 2, True),
                                                  (2, 90, 122, 64, 7, 3, False), (2, 64, 320, 64, 7, 3, False)])
def test_first_layer_tensor_core_conv_is_fp32_accurate(B, H, W, Cout, K, pad, bias, monkeypatch):
    """Cin = 3 first layers (ResNet stem 7x7/s2/p3, EasyBackbone conv1 5x5/s2/p2) of the training forward: the Toeplitz GEMM of
    csrc/stem_tc.cu with split operands (amoe_stem_fwd_f32tc) against fp64 and against the CUDA-core kernel; frames it does not
    take (W/2 > 128) fall back to the CUDA cores; the weight gradient (CUDA cores in both modes) is unchanged."""
    import torch.nn.functional as F
    from automoe_b200.training import functional as TF
    torch.backends.cudnn.allow_tf32 = False
    g = _gen(300 + H + K)
    conv = nn.Conv2d(3, Cout, K, 2, pad, bias=bias).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (3 * K * K)) ** 0.5)
        if bias:
            conv.bias.copy_(torch.randn(Cout, generator=g) * 0.1)
    img = torch.randn((B, 3, H, W), generator=g).to(DEV)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("AMOE_TRAIN_TC", flag)
        conv.requires_grad_(False)                       # frozen first layer (the experts' stems): tensor-core forward
        y_frozen = TF.conv_bn_act(TF.image_nhwc4(img), conv, None, relu=False)
        conv.requires_grad_(True)                        # trained first layer (policy conv1): CUDA-core forward and weight gradient
        conv.weight.grad = None
        y = TF.conv_bn_act(TF.image_nhwc4(img), conv, None, relu=False)
        gy = torch.randn(y.shape, generator=_gen(9)).to(DEV)
        y.backward(gy)
        res[flag] = (y_frozen.detach(), conv.weight.grad.detach().clone(), y.detach())
    monkeypatch.setenv("AMOE_TRAIN_TC", "1")
    assert TF._stem_tc_ok(4, 3, Cout, K, K, 2, pad, H, W) == (W // 2 <= 128)
    assert torch.equal(res["1"][2], res["0"][2]) and torch.equal(res["0"][0], res["0"][2])
    if W // 2 <= 128:
        assert not torch.equal(res["1"][0], res["0"][0])      # the tensor-core path was taken
    w64 = conv.weight.detach().double().requires_grad_(True)
    y64 = F.conv2d(img.double(), w64, conv.bias.detach().double() if bias else None, 2, pad)
    y64.backward(gy.double().permute(0, 3, 1, 2))
    y32 = F.conv2d(img, conv.weight.detach(), conv.bias.detach() if bias else None, 2, pad)
    ref = y64.detach().permute(0, 2, 3, 1)
    e, e0, e32 = rel_err(res["1"][0], ref), rel_err(res["0"][0], ref), rel_err(y32.permute(0, 2, 3, 1), ref)
    print("first layer on the tensor cores vs fp64:", e, " CUDA cores:", e0, " torch fp32:", e32)
    assert e < 1e-5 and e0 < 1e-5, (e, e0, e32)
    assert rel_err(res["1"][1], w64.grad) < 1e-5 and torch.equal(res["1"][1], res["0"][1])


@pytest.mark.parametrize("B,H,W,Cin,Cout,K,stride,pad", [
    (2, 16, 16, 64, 64, 3, 1, 1), (3, 14, 18, 64, 128, 3, 2, 1), (2, 16, 12, 128, 128, 1, 2, 0), (1, 8, 8, 256, 512, 3, 1, 1),
    (2, 8, 8, 512, 256, 3, 1, 1), (2, 23, 40, 128, 32, 3, 1, 1), (2, 10, 10, 64, 64, 1, 1, 0),
    (2, 45, 80, 256, 512, 3, 2, 1), (2, 45, 80, 256, 512, 1, 2, 0), (1, 13, 11, 64, 64, 3, 2, 1)])     # odd sizes at stride 2 (720p layer4)
def test_split_operand_tensor_core_conv_is_fp32_accurate(B, H, W, Cin, Cout, K, stride, pad, monkeypatch):
    """fp32-accurate convolution on the bf16 tensor cores (three-way operand split, six product terms, fp32 TMEM
    accumulation): forward and data gradient against an fp64 evaluation, within 1e-5 relative (max-norm) - the tensor core
    truncates when it adds into the fp32 accumulator, ~2^-25 of the sum per leading-term MMA, which any bf16/TF32 GEMM pays
    too; the CUDA-core path (AMOE_TRAIN_TC=0) gives the same answer to fp32 rounding."""
    import torch.nn.functional as F
    from automoe_b200.training import functional as TF
    torch.backends.cudnn.allow_tf32 = False
    g = _gen(100 + Cin + Cout + K + stride)
    conv = nn.Conv2d(Cin, Cout, K, stride, pad, bias=True).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (Cin * K * K)) ** 0.5)
        conv.bias.copy_(torch.randn(Cout, generator=g) * 0.1)
    x = torch.randn((B, H, W, Cin), generator=g).to(DEV)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("AMOE_TRAIN_TC", flag)
        conv.weight.grad = None
        xi = x.clone().requires_grad_(True)
        y = TF.conv_bn_act(xi, conv, None, relu=False)
        gy = torch.randn(y.shape, generator=_gen(7)).to(DEV)
        y.backward(gy)
        res[flag] = (y.detach(), xi.grad.detach(), conv.weight.grad.detach().clone())
    x64 = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    w64 = conv.weight.detach().double().requires_grad_(True)
    y64 = F.conv2d(x64, w64, conv.bias.detach().double(), stride, pad)
    gy64 = gy.double().permute(0, 3, 1, 2)
    y64.backward(gy64)
    x32 = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    y32 = F.conv2d(x32, conv.weight.detach(), conv.bias.detach(), stride, pad)
    y32.backward(gy.permute(0, 3, 1, 2).contiguous())
    y_ref, dx_ref = y64.detach().permute(0, 2, 3, 1), x64.grad.permute(0, 2, 3, 1)
    for name, ours, ref, torch32 in (("y", res["1"][0], y_ref, y32.detach().permute(0, 2, 3, 1)),
                                     ("dx", res["1"][1], dx_ref, x32.grad.permute(0, 2, 3, 1))):
        e, e32 = rel_err(ours, ref), rel_err(torch32, ref)
        print(name, "tensor-core split conv vs fp64:", e, " torch fp32 vs fp64:", e32)
        assert e < 1e-5, (name, e, e32)
        assert rel_err(res["0"][0 if name == "y" else 1], ref) < 1e-5
    assert rel_err(res["1"][2], w64.grad) < 1e-5            # weight gradient: CUDA-core kernel in both modes
