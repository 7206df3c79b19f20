"""2-GPU check of the data-parallel gating/policy training step (BASELINE.json configs[3]):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/train_ddp_check.py

Each rank takes half of a 16-frame batch, runs forward/backward through the sm_100a training kernels, and
FlatAdamW.step() all-reduces the ONE flat gradient buffer over NCCL (NVLink), clips by the global norm and
applies AdamW in one kernel.  Checks: (1) parameters stay bit-identical across ranks, (2) the averaged
gradient equals the full-batch gradient computed on one GPU (eval-mode semantics so BatchNorm statistics do
not depend on the shard), (3) reports the step time at 32 frames per GPU, 256x256, (4) replays the same step as one CUDA graph
(GraphedTrainStep: NCCL all-reduce and buffer broadcast captured) and checks parameters and buffers stay identical across ranks.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import synth  # noqa: E402  (test infrastructure: seeded weights / inputs only)


def targets(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((B, 10, 2), generator=g) * 5.0, torch.rand((B, 10), generator=g) * 30.0


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    from automoe_b200.models.automoe import create_automoe_model
    from automoe_b200.training.train_gating_network import FlatAdamW, compute_gating_losses, freeze_for_gating_training, train_step

    def build():
        cfg = dict(synth.CONFIG_3EXPERT, precision="fp32")
        m = create_automoe_model(cfg, "cpu")
        m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0))
        m = m.to(dev).eval()
        m.device = dev
        return m

    # ---- (2) averaged shard gradients == full-batch gradient ----
    B = 8 * world
    full = synth.synth_batch(B, 64, 64, seed=3)
    wp, spd = targets(B, 4)
    m = build()
    params = freeze_for_gating_training(m)
    opt = FlatAdamW(params, lr=1e-3, max_norm=1.0)
    sl = slice(rank * 8, rank * 8 + 8)
    shard = {k: v[sl].to(dev) for k, v in full.items()}
    opt.zero_grad()
    losses = compute_gating_losses(m(shard), wp[sl].to(dev), spd[sl].to(dev), {"use_load_balancing": False})
    losses["total_loss"].backward()
    g = opt.flat_grad.clone()
    dist.all_reduce(g)
    g /= world
    m2 = build()
    p2 = freeze_for_gating_training(m2)
    opt2 = FlatAdamW(p2, lr=1e-3, max_norm=1.0)
    opt2.zero_grad()
    fb = {k: v.to(dev) for k, v in full.items()}
    l2 = compute_gating_losses(m2(fb), wp.to(dev), spd.to(dev), {"use_load_balancing": False})
    l2["total_loss"].backward()
    err = ((g - opt2.flat_grad).abs().max() / opt2.flat_grad.abs().max()).item()
    # ---- (1) three optimizer steps: parameters identical on every rank ----
    for _ in range(3):
        opt.zero_grad()
        compute_gating_losses(m(shard), wp[sl].to(dev), spd[sl].to(dev), {})["total_loss"].backward()
        opt.step()
    mine = opt.flat_param.clone()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(mine, ref))
    # ---- (3) step time at the BASELINE shape ----
    m.train()
    b32 = {k: v.to(dev) for k, v in synth.synth_batch(32, 256, 256, seed=10 + rank).items()}
    w32, s32 = targets(32, 11 + rank)
    b32["waypoints"], b32["speed"] = w32.to(dev), s32.to(dev)
    for _ in range(3):
        train_step(m, b32, opt, {})
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        train_step(m, b32, opt, {})
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # ---- (4) the same step captured as one CUDA graph (all-reduce and buffer broadcast inside): replicas stay identical ----
    from automoe_b200.training.train_gating_network import GraphedTrainStep
    import gc
    del losses, l2          # autograd graphs of the eager checks above: their AccumulateGrad nodes sit on the default stream
    gc.collect()
    graphed = GraphedTrainStep(m, b32, opt, {})
    for _ in range(3):
        graphed(b32)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        graphed(b32)
    e1.record()
    torch.cuda.synchronize()
    ms_g = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(ms_g, op=dist.ReduceOp.MAX)
    mine = torch.cat([opt.flat_param] + [b.reshape(-1).float() for b in m.buffers()])
    ref = mine.clone()
    dist.broadcast(ref, 0)
    same_g = bool(torch.equal(mine, ref)) and bool(torch.isfinite(mine).all())
    if rank == 0:
        print({"world": world, "avg_shard_grad_vs_full_batch_rel_err": err, "params_identical_across_ranks": same,
               "train_step_ms_b32_per_gpu": ms.item(), "frames_per_s": world * 32 / (ms.item() / 1e3),
               "graphed_step_ms_b32_per_gpu": ms_g.item(), "graphed_frames_per_s": world * 32 / (ms_g.item() / 1e3),
               "graphed_params_and_buffers_identical_across_ranks": same_g}, flush=True)
    del graphed             # the captured graph holds the NCCL communicator: release it before the process group goes away
    gc.collect()
    torch.cuda.synchronize()
    ok = torch.tensor([int(same and same_g and err < 1e-4)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1 else 1)


if __name__ == "__main__":
    main()
