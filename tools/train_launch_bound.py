#!/usr/bin/env python
"""Is the gating training step launch-bound?  Host time to ENQUEUE one step (no sync) against the device time of the step.
python tools/train_launch_bound.py [gating|gating_fast] [batch]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import bench
    import bench_train as BT
    from automoe_b200.models.automoe import create_automoe_model
    from automoe_b200.training.train_gating_network import FlatAdamW, freeze_for_gating_training, train_step
    mode = sys.argv[1] if len(sys.argv) > 1 else "gating"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
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
    ctxm = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "gating_fast" else torch.autocast("cuda", enabled=False)
    with ctxm:
        for _ in range(5):
            train_step(model, batch, opt, {})
        torch.cuda.synchronize()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = []
        e0.record()
        for _ in range(n):
            t = time.perf_counter()
            train_step(model, batch, opt, {})
            host.append(time.perf_counter() - t)
        e1.record()
        t_sync = time.perf_counter()
        torch.cuda.synchronize()
        tail = time.perf_counter() - t_sync
    host.sort()
    print(f"{mode} B={B}: device {e0.elapsed_time(e1) / n:.2f} ms/step; host enqueue median {host[n // 2] * 1e3:.2f} ms/step "
          f"(min {host[0] * 1e3:.2f}); device work still queued after the last enqueue: {tail * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
