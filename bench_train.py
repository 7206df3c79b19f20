"""Training modes of bench.py (BASELINE.json configs[3] and configs[4]); imported by bench.py, same launch contract.

  --mode train_gating     gating + policy training step with frozen experts (training/train_gating_network.py:92-105),
                          batch 32 frames of 3x256x256 per GPU (train_gating_network.sh:35), fp32 like the reference
                          trainer (no autocast there): zero_grad, forward, compute_gating_losses, backward, ONE NCCL
                          all-reduce of the flat 11.5 MB gradient buffer, fused global-norm clip + AdamW (FlatAdamW).
                          `value` is measured with the REFERENCE's train-mode semantics (model.train(): the frozen
                          experts' BatchNorm on batch statistics, running statistics updated); the same step with
                          frozen experts on running statistics through the bf16 tensor-core inference kernels
                          (AutoMoE.frozen_experts_eval = True under autocast) is reported beside it.
  --mode train_detection  detection-expert forward/backward with the Hungarian matcher
                          (training/train_bdd100k_ddp.py:117-186 + :90-99), batch 64 images of 3x720x1280 per GPU,
                          AdamW + clip 1.0, gradients all-reduced once per step.

Every line also carries `gpu_reference`: the same step in stock torch (oracle port of the reference modules:
cuDNN/cuBLAS + torch.autograd + clip_grad_norm_ + torch.optim.AdamW, gradients all-reduced with
torch.distributed) on the same GPU(s) - the "stock eager + DDP" number.
"""
from __future__ import annotations

import json
import os
import time

import torch

METRIC_G = "automoe_gating_train_frames_per_sec"
METRIC_D = "detection_expert_train_images_per_sec"
GFLOP_GATING_PER_FRAME = 3 * 4.888 + 3 * 0.534 + 3 * 0.004      # frozen experts forward + policy fwd/dgrad/wgrad + MLPs
GFLOP_DET_PER_IMAGE = 3 * 69.13                                  # forward + dgrad + wgrad at 720x1280


def _targets(B, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((B, 10, 2), generator=g) * 5.0).to(dev), (torch.rand((B, 10), generator=g) * 30.0).to(dev)


def _det_batch(B, H, W, n_max, seed):
    """images ~ N(0,1); per image U{1..n_max} boxes (xyxy in [0,1]), labels U{0..9}, padded with -1."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn((B, 3, H, W), generator=g)
    boxes = torch.full((B, n_max, 4), -1.0)
    labels = torch.full((B, n_max), -1, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(1, n_max + 1, (1,), generator=g))
        xy = torch.rand((n, 2), generator=g) * 0.6
        wh = torch.rand((n, 2), generator=g) * 0.35 + 0.03
        boxes[b, :n] = torch.cat([xy, xy + wh], dim=1)
        labels[b, :n] = torch.randint(0, 10, (n,), generator=g)
    return {"image": images, "bboxes": boxes, "labels": labels}


def _timed(fn, steps, warmup, barrier, dev, world, dist):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() / steps


def _allreduce_us(n_floats, dev, world, dist):
    if world == 1:
        return 0.0
    buf = torch.zeros(n_floats, device=dev)
    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 100.0], device=dev)     # us per all-reduce
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def _stock_allreduce(grads, world, dist):
    if world > 1:
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


def run_train(args, dev, world, rank, numa):
    import torch.distributed as dist
    from bench import ClockSampler, model_config, peaks, randomize_norm_stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pk = peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if args.mode == "train_gating":
        line = _gating(args, dev, world, rank, dist, barrier, steps, warmup, sampler, pk, model_config, randomize_norm_stats)
    else:
        line = _detection(args, dev, world, rank, dist, barrier, steps, warmup, sampler, pk)
    if rank == 0:
        line["e2e"]["numa"] = numa
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _gating(args, dev, world, rank, dist, barrier, steps, warmup, sampler, pk, model_config, randomize_norm_stats):
    from automoe_b200 import _cabi
    from automoe_b200.models.automoe import create_automoe_model
    from automoe_b200.training.train_gating_network import FlatAdamW, GraphedTrainStep, freeze_for_gating_training, train_step
    B = args.batch or 32
    torch.manual_seed(0)
    model = create_automoe_model(model_config(), "cpu")
    randomize_norm_stats(model)
    model = model.to(dev)
    model.device = dev
    params = freeze_for_gating_training(model)
    opt = FlatAdamW(params, lr=1e-4, weight_decay=1e-4, max_norm=1.0)
    g = torch.Generator().manual_seed(10 + rank)
    batch = {"image": torch.randn((B, 3, 256, 256), generator=g).to(dev), "speed": (torch.rand((B, 1), generator=g) * 30.0).to(dev),
             "steering": torch.zeros(B, 1, device=dev), "throttle": torch.zeros(B, 1, device=dev), "brake": torch.zeros(B, 1, device=dev)}
    wp, spd = _targets(B, 20 + rank, dev)
    # the reference's loader puts the targets under batch["waypoints"] / batch["speed"] (train_gating_network.py:98); the
    # forward then reads batch["speed"][:, -1:] as the vehicle speed (automoe.py:108-111) - same here
    tbatch = dict(batch, waypoints=wp, speed=spd)
    model.train()

    def step_ref():
        return train_step(model, tbatch, opt, {})

    # eager launches from Python first (what round 2 reported), then the same step captured once and replayed
    ms_eager = _timed(step_ref, steps, warmup, barrier, dev, world, dist)
    graphed = GraphedTrainStep(model, tbatch, opt, {})

    def step_graph():
        return graphed(tbatch)

    n0 = _cabi.launch_count(dev)
    sampler.mark_begin()
    ms_ref = _timed(step_graph, steps, warmup, barrier, dev, world, dist)
    sampler.mark_end()
    assert _cabi.launch_count(dev) == n0      # replays launch nothing from the host
    launches = graphed.launches_per_replay * steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms_ref / 1e3)

    # end to end: host (pinned) batch in, loss out, reference semantics, through the same graph
    pinned = {k: v.cpu().pin_memory() for k, v in tbatch.items()}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    loss_host = torch.empty(1).pin_memory()

    def step_e2e():
        out = graphed(pinned)
        loss_host.copy_(out["total_loss"].reshape(1), non_blocking=True)
    ms_e2e = _timed(step_e2e, steps, 2, barrier, dev, world, dist)
    del graphed

    # the opt-out: frozen experts on running statistics (inference kernels), fp32 and under bf16 autocast
    model.frozen_experts_eval = True
    ms_fast32 = _timed(step_ref, steps, warmup, barrier, dev, world, dist)
    g32 = GraphedTrainStep(model, tbatch, opt, {})
    ms_fast32_graph = _timed(lambda: g32(tbatch), steps, warmup, barrier, dev, world, dist)
    del g32

    def step_fast16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return train_step(model, tbatch, opt, {})
    ms_fast16 = _timed(step_fast16, steps, warmup, barrier, dev, world, dist)
    g16 = GraphedTrainStep(model, tbatch, opt, {}, autocast_dtype=torch.bfloat16)
    ms_fast16_graph = _timed(lambda: g16(tbatch), steps, warmup, barrier, dev, world, dist)
    del g16
    model.frozen_experts_eval = False
    ar_us = _allreduce_us(opt.n, dev, world, dist)

    # stock torch on the same GPU(s): oracle port of the reference modules + autograd + clip + AdamW (+ all-reduce)
    gref = None
    if not args.no_gpu_reference:
        gref = _stock_gating(B, dev, world, rank, dist, barrier, steps, warmup, model_config)
        gref["ours_over_gpu_reference"] = value / gref["value"]
        # the stock step keeps the frozen experts on running statistics: the like-for-like comparison is our same-semantics variant
        gref["ours_same_semantics_over_gpu_reference"] = gref["ms_per_step"] / ms_fast32_graph
    if rank != 0:
        return None
    ach = value / world * GFLOP_GATING_PER_FRAME / 1e3
    return {
        "metric": METRIC_G, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_ref, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Gating + policy training step (frozen experts) with NCCL grad allreduce (BASELINE.json configs[3])",
                   "batch_per_gpu": B, "image": "3x256x256", "optimizer": "FlatAdamW (one all-reduce, fused clip 1.0 + AdamW)",
                   "semantics": "reference train mode: model.train(), frozen experts on batch statistics (train_gating_network.py:85)",
                   "parallelism": f"data-parallel x{world}, one gradient all-reduce per step",
                   "launch_mode": "cuda_graph replay of GraphedTrainStep (whole step: forward, losses, backward, all-reduce, clip + AdamW; "
                                  "step count and dropout key on the device); eager launches reported under variants",
                   "l2_policy": "activations of a 32-frame fp32 step (~3 GB) exceed the 126 MB L2"},
        "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clocks,
        "allreduce": {"bytes": opt.n * 4, "us": ar_us, "share_of_step": ar_us / 1e3 / ms_ref},
        "variants": {"reference_semantics_fp32": {"ms_per_step": ms_ref, "frames_per_s": world * B / (ms_ref / 1e3),
                                                  "eager_ms_per_step": ms_eager},
                     "frozen_experts_eval_fp32": {"ms_per_step": ms_fast32_graph, "frames_per_s": world * B / (ms_fast32_graph / 1e3),
                                                  "eager_ms_per_step": ms_fast32},
                     "frozen_experts_eval_bf16_autocast": {"ms_per_step": ms_fast16_graph, "frames_per_s": world * B / (ms_fast16_graph / 1e3),
                                                           "eager_ms_per_step": ms_fast16,
                                                           "note": "frozen experts through the tcgen05 inference kernels; trainable part fp32"}},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                     "traffic": None, "kernel": "fp32-accurate split-operand tcgen05 convolutions (conv_tc_kernel: six bf16 MMAs per fp32 MAC, so the bf16 peak overstates the reachable rate 6x) + CUDA-core remainder (policy first layers, stride-2 weight gradients, small-batch MLP GEMMs)",
                     "algorithmic_gflop_per_frame": GFLOP_GATING_PER_FRAME},
        "cpu_baseline": None, "gpu_reference": gref,
    }


def _stock_gating(B, dev, world, rank, dist, barrier, steps, warmup, model_config):
    from oracle import gating_train_oracle as GT
    from oracle import synth
    from automoe_b200.models.automoe import create_automoe_model
    cfg = model_config()
    sd = synth.synth_state_dict(create_automoe_model(cfg, "cpu").state_dict(), 0)
    sdd = {k: v.to(dev).clone() for k, v in sd.items()}
    train = [v.requires_grad_(True) for k, v in sdd.items() if GT.is_trainable_key(k) and v.is_floating_point()]
    opt = torch.optim.AdamW(train, lr=1e-4, weight_decay=1e-4)
    g = torch.Generator().manual_seed(10 + rank)
    batch = {"image": torch.randn((B, 3, 256, 256), generator=g).to(dev), "speed": (torch.rand((B, 1), generator=g) * 30.0).to(dev),
             "steering": torch.zeros(B, 1, device=dev), "throttle": torch.zeros(B, 1, device=dev), "brake": torch.zeros(B, 1, device=dev)}
    wp, spd = _targets(B, 20 + rank, dev)
    torch.backends.cudnn.benchmark = True

    def step():
        opt.zero_grad(set_to_none=False)
        pred = GT.training_forward(sdd, batch, cfg, policy_batch_stats=True)
        GT.compute_gating_losses(pred, wp, spd, {})["total_loss"].backward()
        _stock_allreduce([p.grad for p in train if p.grad is not None], world, dist)
        torch.nn.utils.clip_grad_norm_(train, max_norm=1.0)
        opt.step()
    ms = _timed(step, steps, warmup, barrier, dev, world, dist)
    torch.backends.cudnn.benchmark = False
    return {"value": world * B / (ms / 1e3), "unit": "frames/s", "ms_per_step": ms,
            "what": "oracle port of the reference modules in stock torch (cuDNN/cuBLAS fp32, TF32 off, autograd, clip_grad_norm_, "
                    "torch.optim.AdamW, flat gradient all-reduce), frozen experts on running statistics, cudnn.benchmark=True"}


def _detection(args, dev, world, rank, dist, barrier, steps, warmup, sampler, pk):
    from automoe_b200 import _cabi
    from automoe_b200.models.experts import BDDDetectionExpert
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
    from automoe_b200.training.train_bdd100k import train_detection_batch
    from automoe_b200.training.train_gating_network import FlatAdamW
    B = args.batch or 64
    H, W = (int(v) for v in args.det_hw.lower().split("x"))
    torch.manual_seed(0)
    model = BDDDetectionExpert(num_classes=10, pretrained_backbone=False).to(dev).train()
    opt = FlatAdamW(list(model.parameters()), lr=1e-4, weight_decay=1e-4, max_norm=1.0)
    matcher = HungarianMatcher()
    hb = _det_batch(B, H, W, 60, 30 + rank)
    batch = {k: v.to(dev) for k, v in hb.items()}

    def step(b=batch):
        opt.zero_grad()
        loss = train_detection_batch(model, b, matcher)
        loss.backward()
        opt.step()
        return loss

    n0 = _cabi.launch_count(dev)
    sampler.mark_begin()
    ms = _timed(step, steps, warmup, barrier, dev, world, dist)
    sampler.mark_end()
    launches = (_cabi.launch_count(dev) - n0) * steps // (steps + warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms / 1e3)
    h, w = H, W
    for _ in range(5):                      # conv1, max-pool, layer2-4: five stride-2 stages
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    Q = h * w
    pinned = {k: v.pin_memory() for k, v in hb.items()}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    loss_host = torch.empty(1).pin_memory()

    def step_e2e():
        b = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        loss_host.copy_(step(b).detach().reshape(1), non_blocking=True)
    ms_e2e = _timed(step_e2e, max(2, steps // 2), 1, barrier, dev, world, dist)
    ar_us = _allreduce_us(opt.n, dev, world, dist)
    gref = None
    if not args.no_gpu_reference:
        del opt
        torch.cuda.empty_cache()
        gref = _stock_detection(B, H, W, dev, world, rank, dist, barrier, max(2, steps // 2), 2)
        gref["ours_over_gpu_reference"] = value / gref["value"]
    if rank != 0:
        return None
    ach = value / world * GFLOP_DET_PER_IMAGE * (H * W) / (720 * 1280) / 1e3
    return {
        "metric": METRIC_D, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Detection-expert training forward/backward with Hungarian matcher cost matrix (BASELINE.json configs[4])",
                   "batch_per_gpu": B, "image": f"3x{H}x{W}", "queries": Q, "targets_per_image": "U{1..60}",
                   "optimizer": "FlatAdamW (one all-reduce of 49.4 MB, fused clip 1.0 + AdamW)",
                   "parallelism": f"data-parallel x{world}, one gradient all-reduce per step",
                   "l2_policy": "activations (GBs per step) exceed the 126 MB L2"},
        "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clocks,
        "allreduce": {"bytes": 12360014 * 4, "us": ar_us, "share_of_step": ar_us / 1e3 / ms},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                     "traffic": None, "kernel": "fp32-accurate split-operand tcgen05 convolutions (conv_tc_kernel fwd / dgrad, wgrad_tc_kernel: six bf16 MMAs per fp32 MAC, so the bf16 peak overstates the reachable rate 6x) + CUDA-core remainder (720p stem, stride-2 weight gradients, odd-height stride-2 layers)",
                     "algorithmic_gflop_per_image": GFLOP_DET_PER_IMAGE * (H * W) / (720 * 1280)},
        "cpu_baseline": None, "gpu_reference": gref,
    }


def _stock_detection(B, H, W, dev, world, rank, dist, barrier, steps, warmup):
    from oracle import detection_train_oracle as DO
    from oracle import synth
    from automoe_b200.models.automoe import create_automoe_model
    from bench import model_config
    full = synth.synth_state_dict(create_automoe_model(model_config(), "cpu").state_dict(), 0)
    sd = {k[len("experts.0."):]: v.to(dev).clone() for k, v in full.items() if k.startswith("experts.0.")}
    train = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    opt = torch.optim.AdamW(train, lr=1e-4, weight_decay=1e-4)
    batch = {k: v.to(dev) for k, v in _det_batch(B, H, W, 60, 30 + rank).items()}
    torch.backends.cudnn.benchmark = True

    def step():
        opt.zero_grad(set_to_none=False)
        out = DO.detection_forward_train(batch["image"], sd, 10, batch_stats=True)
        DO.detection_loss(out, batch["bboxes"], batch["labels"])["total_loss"].backward()
        _stock_allreduce([p.grad for p in train if p.grad is not None], world, dist)
        torch.nn.utils.clip_grad_norm_(train, max_norm=1.0)
        opt.step()
    ms = _timed(step, steps, warmup, barrier, dev, world, dist)
    torch.backends.cudnn.benchmark = False
    return {"value": world * B / (ms / 1e3), "unit": "images/s", "ms_per_step": ms,
            "what": "oracle port of _train_detection_batch in stock torch (cuDNN fp32, TF32 off, autograd, per-image scipy matcher, "
                    "clip_grad_norm_, torch.optim.AdamW, flat gradient all-reduce), cudnn.benchmark=True"}


def run_reference_train(args):
    """--impl reference for the training modes: the oracle port on the host cores, bounded sample."""
    from bench import model_config
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    if args.mode == "train_gating":
        from oracle import gating_train_oracle as GT
        from oracle import synth
        from automoe_b200.models.automoe import create_automoe_model
        cfg = model_config()
        sd = synth.synth_state_dict(create_automoe_model(cfg, "cpu").state_dict(), 0)
        train = [v.requires_grad_(True) for k, v in sd.items() if GT.is_trainable_key(k) and v.is_floating_point()]
        opt = torch.optim.AdamW(train, lr=1e-4, weight_decay=1e-4)
        B = 8
        g = torch.Generator().manual_seed(10)
        batch = {"image": torch.randn((B, 3, 256, 256), generator=g), "speed": torch.rand((B, 1), generator=g) * 30.0,
                 "steering": torch.zeros(B, 1), "throttle": torch.zeros(B, 1), "brake": torch.zeros(B, 1)}
        wp, spd = _targets(B, 20, "cpu")

        def step():
            opt.zero_grad()
            pred = GT.training_forward(sd, batch, cfg, policy_batch_stats=True)
            GT.compute_gating_losses(pred, wp, spd, {})["total_loss"].backward()
            torch.nn.utils.clip_grad_norm_(train, max_norm=1.0)
            opt.step()
        metric, unit, sample = METRIC_G, "frames/s", f"{B} frames of 3x256x256 per step, fp32 oracle port (frozen experts on running statistics)"
    else:
        from oracle import detection_train_oracle as DO
        from oracle import synth
        from automoe_b200.models.automoe import create_automoe_model
        full = synth.synth_state_dict(create_automoe_model(model_config(), "cpu").state_dict(), 0)
        sd = {k[len("experts.0."):]: v.clone() for k, v in full.items() if k.startswith("experts.0.")}
        train = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
        opt = torch.optim.AdamW(train, lr=1e-4, weight_decay=1e-4)
        B = 1
        batch = _det_batch(B, 720, 1280, 60, 30)

        def step():
            opt.zero_grad()
            out = DO.detection_forward_train(batch["image"], sd, 10, batch_stats=True)
            DO.detection_loss(out, batch["bboxes"], batch["labels"])["total_loss"].backward()
            torch.nn.utils.clip_grad_norm_(train, max_norm=1.0)
            opt.step()
        metric, unit, sample = METRIC_D, "images/s", f"{B} image of 3x720x1280 per step, fp32 oracle port of _train_detection_batch"
    steps = min(args.steps, 5)
    for _ in range(1):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    per = (time.perf_counter() - t0) / steps
    v = B / per
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
        "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"bench.py --mode {args.mode}", "sample": sample},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
