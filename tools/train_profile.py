#!/usr/bin/env python
"""Top kernels of one training step (torch.profiler, CUDA time): python tools/train_profile.py [gating|gating_fast|detection] [batch]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import bench
    import bench_train as BT
    argv = [a for a in sys.argv if not a.startswith("--")]
    mode = argv[1] if len(argv) > 1 else "gating"
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    if mode.startswith("gating"):
        from automoe_b200.models.automoe import create_automoe_model
        from automoe_b200.training.train_gating_network import FlatAdamW, freeze_for_gating_training, train_step
        B = int(argv[2]) if len(argv) > 2 else 32
        model = create_automoe_model(bench.model_config(), "cpu")
        bench.randomize_norm_stats(model)
        model = model.to(dev)
        model.device = dev
        opt = FlatAdamW(freeze_for_gating_training(model), lr=1e-4, weight_decay=1e-4, max_norm=1.0)
        g = torch.Generator().manual_seed(10)
        batch = {"image": torch.randn((B, 3, 256, 256), generator=g).to(dev), "steering": torch.zeros(B, 1, device=dev),
                 "throttle": torch.zeros(B, 1, device=dev), "brake": torch.zeros(B, 1, device=dev)}
        wp, spd = BT._targets(B, 20, dev)
        batch.update(waypoints=wp, speed=spd)
        model.train()
        model.frozen_experts_eval = mode == "gating_fast"
        step = lambda: train_step(model, batch, opt, {})
    else:
        from automoe_b200.models.experts import BDDDetectionExpert
        from automoe_b200.training.hungarian_matcher import HungarianMatcher
        from automoe_b200.training.train_bdd100k import train_detection_batch
        from automoe_b200.training.train_gating_network import FlatAdamW
        B = int(argv[2]) if len(argv) > 2 else 8
        model = BDDDetectionExpert(num_classes=10, pretrained_backbone=False).to(dev).train()
        opt = FlatAdamW(list(model.parameters()), lr=1e-4, weight_decay=1e-4, max_norm=1.0)
        matcher = HungarianMatcher()
        batch = {k: v.to(dev) for k, v in BT._det_batch(B, 720, 1280, 60, 30).items()}

        def step():
            opt.zero_grad()
            train_detection_batch(model, batch, matcher).backward()
            opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            t = rows.setdefault(e.name[:110], [0.0, 0])
            t[0] += e.device_time
            t[1] += 1
    detail = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and any(k in e.name for k in ("conv2d_simt", "conv_bwd_weight", "conv_bwd_data")):
            detail.setdefault(e.name[:60], []).append(round(e.device_time))
    for name, ts in detail.items():      # per-launch times (us) of the CUDA-core convolutions, in launch order, last profiled step
        print(f"#   {name}: {ts[-(len(ts) // 3):]}")
    if "--trace" in sys.argv:            # every kernel of the last profiled step in launch order (us)
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        evs = evs[-(len(evs) // 3):]
        for e in evs:
            print(f"T {e.device_time:8.1f}  {e.name[:90]}")
    tot = sum(v[0] for v in rows.values())
    print(f"# {mode}: {tot / 3 / 1e3:.2f} ms of kernel time per step")
    for name, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"{t / 3 / 1e3:9.3f} ms  {n // 3:5d}x  {name}")


if __name__ == "__main__":
    main()
