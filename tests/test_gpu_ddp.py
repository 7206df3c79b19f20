"""GPU (-m gpu, needs >= 2 devices): the module is wrappable by torch.nn.parallel.DistributedDataParallel exactly as the
reference trainer does (training/train_gating_network.py:235-236: `model = DDP(model, device_ids=[rank])`), one process
per GPU over NCCL: one reference-style step (zero_grad, forward, compute_gating_losses, backward, clip_grad_norm_ 1.0,
AdamW) leaves identical parameters on both ranks, and DDP's averaged gradient equals the full-batch gradient."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    for p in (str(root), str(root / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from _util import build_b200_model
    from oracle import synth
    from automoe_b200.training.train_gating_network import compute_gating_losses
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    dev = f"cuda:{rank}"
    m, _ = build_b200_model(dev, "fp32")
    m.freeze_experts()
    m.train()
    m.experts.eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ddp = DDP(m, device_ids=[rank])                      # the reference's wrapping
    opt = torch.optim.AdamW([p for p in ddp.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-4)
    B = 8
    full = synth.synth_batch(B, 64, 64, seed=3, speed_seq=1)
    g = torch.Generator().manual_seed(4)
    wp, spd = torch.randn((B, 10, 2), generator=g) * 5.0, torch.rand((B, 10), generator=g) * 30.0
    sl = slice(rank * B // world, (rank + 1) * B // world)
    shard = {k: v[sl].to(dev) for k, v in full.items()}
    opt.zero_grad()
    pred = ddp(shard)
    losses = compute_gating_losses(pred, wp[sl].to(dev), spd[sl].to(dev), {})
    losses["total_loss"].backward()                      # DDP all-reduces (averages) the gradients here
    name = "gating_network.gate_network.3.weight"
    g_ddp = dict(m.named_parameters())[name].grad.clone()
    torch.nn.utils.clip_grad_norm_(ddp.parameters(), max_norm=1.0)
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters() if p.requires_grad])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(flat, ref))
    # full-batch gradient on one device (no DDP): mean over the batch == mean of the two shard means
    m2, _ = build_b200_model(dev, "fp32")
    m2.freeze_experts()
    m2.train()
    m2.experts.eval()
    for mod in m2.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    # batch-statistics BatchNorm of the policy backbone sees the shard, not the full batch: compare a head-only gradient path
    pred2 = m2({k: v.to(dev) for k, v in full.items()})
    loss2 = compute_gating_losses(pred2, wp.to(dev), spd.to(dev), {})
    loss2["total_loss"].backward()
    g_full = dict(m2.named_parameters())[name].grad
    err = ((g_ddp - g_full).abs().max() / g_full.abs().max()).item()
    if rank == 0:
        out.put({"same_params": same, "grad_err_vs_full_batch": err, "loss": float(losses["total_loss"])})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_wraps_in_ddp_and_steps_like_the_reference_trainer():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = out.get(timeout=10)
    print("DDP world-2:", res)
    assert res["same_params"]
    # load-balancing / entropy terms are batch-mean based (not linear in the shard split) and the policy BatchNorm uses shard
    # statistics: the averaged shard gradient of the gate output layer agrees with the full-batch one to a few percent
    assert res["grad_err_vs_full_batch"] < 0.2
