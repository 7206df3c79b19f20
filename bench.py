#!/usr/bin/env python
"""bench.py — AutoMoE batched forward, frames/s (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward pass of the AutoMoE hot path (3 experts + gating + policy) over one
batch of 256 synthetic 3x256x256 frames per GPU (BASELINE.json configs[1]); weights are
random-init (seeded), BatchNorm statistics perturbed.  Frames are independent, so N GPUs
run N independent shards with no data-path collective (weak scaling).

  value  : whole-job frames/s with inputs resident in HBM (device-timed, max over ranks)
  e2e    : frames/s through the public module API with HOST (pinned) inputs: every step
           uploads its fp32 frames + vehicle state and downloads waypoints/speed/weights
  roofline: tensor bound; achieved = algorithmic conv FLOPs of the tcgen05 conv launches /
           their CUDA-event time, measured live after the timed region
  cpu_baseline / --impl reference: the oracle port of the reference (pure torch CPU fp32,
           all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GFLOP_PER_FRAME = 15.203  # SURVEY.md §8 / BASELINE.md §3 (2*MAC over conv+linear, counted on the reference)
METRIC = "automoe_forward_frames_per_sec"
UNIT = "frames/s"
WORKLOAD = "AutoMoE batched inference bf16, batch=256 synthetic 3x256x256 frames per GPU (BASELINE.json configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=16, help="frames per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying AutoMoE.capture()")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def conv_traffic(launches_per_step):
    """DRAM bytes per launch of the tensor-core conv launches, from the committed ncu pass (tools/summarize_launches.py);
    None when the capture is missing or was taken with a different launch count."""
    p = ROOT / "profiles" / "conv_traffic.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text())
    if d.get("conv_launches") != launches_per_step:
        return None
    return d["dram_bytes_per_launch"]


def model_config():
    return {
        "experts": [
            {"type": "detection", "num_classes": 10, "output_dim": 256, "pretrained_backbone": False},
            {"type": "segmentation", "num_classes": 19, "output_dim": 256, "pretrained_backbone": False},
            {"type": "drivable", "num_classes": 3, "output_dim": 256, "pretrained_backbone": False},
        ],
        "gating": {"processed_dim": 256, "hidden_dim": 128, "temperature": 1.0, "use_softmax": True},
        "context": {"type": "simple", "context_dim": 64},
        "policy": {"hidden_dim": 256, "num_waypoints": 10, "waypoint_dim": 2},
    }


def randomize_norm_stats(model, seed=0):
    """BN running stats / affine away from identity so the folded epilogue is exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def host_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    return {
        "image": torch.randn((B, 3, 256, 256), generator=g),
        "speed": torch.rand((B, 1), generator=g) * 30.0,
        "steering": torch.zeros(B, 1), "throttle": torch.zeros(B, 1), "brake": torch.zeros(B, 1),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # median over the upper half = samples taken under load
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_throughput(sample, steps, warmup):
    """The oracle port of the reference forward on the host cores (fp32, all threads)."""
    from oracle import automoe_oracle as O
    from oracle import synth
    from automoe_b200.models.automoe import create_automoe_model
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = model_config()
    template = create_automoe_model(cfg, "cpu").state_dict()
    sd = synth.synth_state_dict(template, 0)
    batch = host_batch(sample, 1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.automoe_forward(sd, batch, cfg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return sample * len(times) / total, cores, total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, cores, per_step = cpu_reference_throughput(args.cpu_sample, args.steps, args.warmup)
    sample = f"{args.cpu_sample} frames of 3x256x256 per step, fp32, oracle port of the reference forward (torch CPU ops)"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")

    from automoe_b200 import _cabi, _ops
    from automoe_b200.models.automoe import create_automoe_model

    torch.manual_seed(0)
    model = create_automoe_model(model_config(), "cpu")
    randomize_norm_stats(model)
    model = model.to(dev).eval()
    model.device = dev
    B = args.batch
    hb = host_batch(B, 1 + rank)
    batch = {k: v.to(dev) for k, v in hb.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def eager_step(b):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(b)

    use_graph = not args.no_graph and os.environ.get("AMOE_GRAPH", "1") != "0"
    graphs = {}

    def step(b):
        """One forward through the public API: AutoMoE.capture() replay (one graph per input buffer set,
        captured in place on first use) or, with --no-graph, the eager module call."""
        if not use_graph:
            return eager_step(b)
        g = graphs.get(id(b))
        if g is None:
            eager_step(b)                      # pack weights before capture
            torch.cuda.synchronize()
            g = graphs[id(b)] = model.capture(b, clone_inputs=False)
        return g()

    # ---------------- device-resident throughput (value) ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # nvidia-smi needs ~100 ms to start: launch it before the warm-up
    for _ in range(max(args.warmup, 3)):
        out = step(batch)
    barrier()
    n0 = _cabi.launch_count(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step(batch)
    e1.record()
    torch.cuda.synchronize()
    launches = _cabi.launch_count(dev) - n0
    if use_graph:
        launches = graphs[id(batch)].launches_per_replay * args.steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        # short timed regions end before nvidia-smi has produced samples: keep the same load running
        # (untimed) until a handful of samples under load exist
        t_wait = time.time()
        while len(sampler.lines) < 8 and time.time() - t_wait < 3.0:
            step(batch)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms = ms.item()
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- end-to-end through the module API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        pinned = [{k: v.pin_memory() for k, v in host_batch(B, 100 + rank * 7 + i).items()} for i in range(2)]
        h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
        res_host = [torch.empty((B, 10, 2)).pin_memory(), torch.empty((B, 1)).pin_memory(), torch.empty((B, 3)).pin_memory()]
        d2h = sum(t.numel() * t.element_size() for t in res_host)
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        slots = [{k: torch.empty_like(v, device=dev) for k, v in pinned[i].items()} for i in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(i):
            s = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                for k, v in pinned[s].items():
                    slots[s][k].copy_(v, non_blocking=True)
                ready[s].record(copy_stream)

        def e2e_loop(n):
            for s in range(2):
                freed[s].record(main)
            upload(0)
            for i in range(n):
                s = i & 1
                if i + 1 < n:
                    upload(i + 1)          # prefetch the next batch while this one computes
                main.wait_event(ready[s])
                o = step(slots[s])
                res_host[0].copy_(o["waypoints"], non_blocking=True)
                res_host[1].copy_(o["speed"], non_blocking=True)
                res_host[2].copy_(o["expert_weights"], non_blocking=True)
                freed[s].record(main)

        e2e_loop(2)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(args.steps)
        t1.record()
        torch.cuda.synchronize()
        ems = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        # raw host->device rate of this box for the same pinned buffer (context for the e2e number)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        slots[0]["image"].copy_(pinned[0]["image"], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = pinned[0]["image"].numel() * 4 / (c0.elapsed_time(c1) / 1e3) / 1e9
        e2e = {"value": world * B * args.steps / (ems.item() / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ems.item() / args.steps,
               "h2d_gbs_measured": h2d_gbs}

    # ---------------- roofline of the dominant kernel (tcgen05 conv), measured live ----------------
    pk = peaks()
    roof = None
    _ops.PROFILE = []
    for _ in range(2):
        eager_step(batch)                      # per-kernel events need the eager launches
    torch.cuda.synchronize()
    rec = _ops.PROFILE
    _ops.PROFILE = None
    agg = {}
    for kind, flops, a, b in rec:
        t, f, n = agg.get(kind, (0.0, 0.0, 0))
        agg[kind] = (t + a.elapsed_time(b), f + flops, n + 1)
    if "conv_tc" in agg:
        t, f, n = agg["conv_tc"]
        ach = f / (t / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all layers)",
                "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "traffic": conv_traffic(n // 2), "peak_source": f"{pk['src']} (MEASURED_PEAKS.json bf16_tflops_sustained)",
                "achieved_per_launch_gflop": f / n / 1e9, "algorithmic_dram_bytes_per_launch": 43e6 * B / (n // 2),
                "launches_per_step": n // 2, "ms_per_step": t / 2,
                "share_of_step": (t / 2) / (ms / args.steps),
                "whole_forward_frac": (value / world) * GFLOP_PER_FRAME * 1e9 / 1e12 / pk["tf_sustained"]}
        if "conv_simt" in agg:
            ts, fs, ns = agg["conv_simt"]
            roof["simt_conv_ms_per_step"] = ts / 2
            roof["simt_conv_tflops"] = fs / (ts / 1e3) / 1e12

    # ---------------- the HBM-bound kernel of the path: the x32 full-resolution logit writer, timed alone ----------------
    # (22 classes x 65536 pixels x 2 B per frame; in the forward it runs on the side stream beside the
    # gate / policy tail, so it is timed here on its own: algorithmic bytes written / CUDA-event time)
    if roof is not None:
        from automoe_b200._cabi import check, ctx as _ctx, dtype_code, lib, ptr, stream_ptr
        low = torch.randn((B, 8, 8, 19), device=dev, dtype=torch.float32)
        up = torch.empty((B, 19, 256, 256), device=dev, dtype=torch.bfloat16)      # written in place by every launch

        def writer():
            check(lib().amoe_upsample_bilinear_nchw_fwd(_ctx(dev), ptr(low), ptr(up), B, 8, 8, 19, 256, 256,
                                                        dtype_code(torch.bfloat16), stream_ptr(dev)), "upsample")
        for _ in range(3):
            writer()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 10                              # back to back: 6.4 GB written, the 126 MB L2 cannot absorb it
        ev[0].record()
        for _ in range(reps):
            writer()
        ev[1].record()
        torch.cuda.synchronize()
        wbytes = up.numel() * 2
        gbs = wbytes * reps / (ev[0].elapsed_time(ev[1]) / 1e3) / 1e9
        roof["hbm_kernel"] = {"kernel": "upsample_intscale_nchw_kernel (x32 bilinear logit writer, 19 classes)", "bound": "hbm",
                              "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                              "algorithmic_bytes_per_launch": wbytes, "us_per_launch": ev[0].elapsed_time(ev[1]) * 1e3 / reps,
                              "how": "10 launches back to back on a preallocated output, CUDA events; the peak is the measured COPY "
                                     "bandwidth (read+write mix) - a pure write stream can exceed it (HBM3e nominal ~7.7 TB/s)"}
        del up, low

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, cores, per = cpu_reference_throughput(args.cpu_sample, 3, 1)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 steps of {args.cpu_sample} frames 3x256x256, fp32 oracle port, {per:.2f} s/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "image": "3x256x256", "experts": 3,
                       "parallelism": f"batch-sharded x{world}, no collective",
                       "l2_policy": "inputs (201 MB fp32 frames per step) and activations exceed the 126 MB L2",
                       "launch_mode": "cuda_graph replay of AutoMoE.capture()" if use_graph else "eager (one ctypes launch per kernel)",
                       "l2_chunk_images": _ops.l2_chunk_images(), "side_stream_outputs": _ops.overlap_outputs()},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
