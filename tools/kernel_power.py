#!/usr/bin/env python
"""Energy per kernel of the AutoMoE forward on a power-capped B200 (runs on the GPU box).

The sustained bench regime is bound by the 1000 W power cap (SM clock floats at ~1.55-1.65 GHz), so what a kernel costs
in a long run is its ENERGY, not its time at boost clocks.  This tool records every kernel-launching call of one eager
forward (batch 256), then replays each call alone in a loop for ~1.2 s while sampling NVML power and SM clock, and prints
per call: time per launch in that steady state, average power, energy per launch (J) and energy per useful TFLOP.

    python tools/kernel_power.py > gpurun_out/kernel_power.csv
"""
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


class Sampler:
    def __init__(self, idx=0):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.s, self.stop = [], False
        self.t = threading.Thread(target=self.run, daemon=True)
        self.t.start()

    def run(self):
        while not self.stop:
            self.s.append((time.perf_counter(), self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                           self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            time.sleep(0.005)

    def window(self, t0, t1):
        w = [x for x in self.s if t0 <= x[0] <= t1]
        if not w:
            return float("nan"), float("nan")
        return sum(x[1] for x in w) / len(w), sum(x[2] for x in w) / len(w)


def main():
    import bench
    from automoe_b200 import _ops
    from automoe_b200.models.automoe import create_automoe_model
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = create_automoe_model(bench.model_config(), "cpu")
    bench.randomize_norm_stats(model)
    model = model.to(dev).eval()
    model.device = dev
    batch = {k: v.to(dev) for k, v in bench.host_batch(256, 1).items()}
    calls = []
    names = ["stage_image_stem", "stem_pool_forward", "conv3x3_flat", "conv2d", "conv2d_dual", "head1x1_pool", "upsample_bilinear_nchw",
             "gate", "policy_head", "mean_hw_nhwc"]
    orig = {n: getattr(_ops, n) for n in names}

    def wrap(n):
        def f(*a, **k):
            calls.append((n, a, k))
            return orig[n](*a, **k)
        return f
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        model(batch)                       # pack weights
        for n in names:
            setattr(_ops, n, wrap(n))
        import automoe_b200.models.experts._trunk as T
        _ops.PROFILE = []
        model(batch)
        prof, _ops.PROFILE = _ops.PROFILE, None
        for n in names:
            setattr(_ops, n, orig[n])
    torch.cuda.synchronize()
    flops = {}
    # PROFILE entries come in call order for the conv-type calls
    conv_calls = [i for i, c in enumerate(calls) if c[0] in ("stem_pool_forward", "conv3x3_flat", "conv2d", "conv2d_dual")]
    for i, (kind, fl, a, b) in zip(conv_calls, prof):
        flops[i] = fl
    smp = Sampler(0)
    # heat up with the whole forward first
    t_end = time.perf_counter() + 2.0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        while time.perf_counter() < t_end:
            model(batch)
        torch.cuda.synchronize()
    print("idx,call,us_per_launch,avg_power_w,avg_sm_mhz,joule_per_launch,gflop,pj_per_flop,detail")
    tot_e = tot_t = 0.0
    for i, (n, a, k) in enumerate(calls):
        f = orig[n]
        with torch.no_grad():
            for _ in range(3):
                f(*a, **k)
            torch.cuda.synchronize()
            reps, t0 = 0, time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            while time.perf_counter() - t0 < 1.2:
                for _ in range(20):
                    f(*a, **k)
                reps += 20
                torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        us = e0.elapsed_time(e1) * 1e3 / reps
        pw, mhz = smp.window(t0 + 0.4, t1)          # skip the ramp
        j = pw * us * 1e-6
        gf = flops.get(i, 0.0) / 1e9
        detail = ""
        if n in ("conv3x3_flat", "conv2d", "conv2d_dual"):
            pc = a[0]
            detail = f"cin={pc.cin} cout={pc.cout} k={pc.kh}x{pc.kw} s={pc.sh} res={k.get('residual') is not None}"
        print(f"{i},{n},{us:.1f},{pw:.0f},{mhz:.0f},{j:.4f},{gf:.1f},{(j / (gf * 1e9) * 1e12) if gf else 0:.3f},{detail}", flush=True)
        tot_e += j
        tot_t += us
    print(f"# sum over calls: {tot_t:.0f} us, {tot_e:.3f} J per forward (each call measured alone in steady state)")
    smp.stop = True


if __name__ == "__main__":
    main()
