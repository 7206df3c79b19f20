"""Batch sharding for multi-GPU inference: frames are independent, so the AutoMoE forward
shards by batch across ranks with a full weight replica per GPU and NO data-path collective
(SURVEY.md §8e).  The only collectives are the timing barrier / max-reduce of a benchmark."""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n_frames over `world` ranks: (start, count); the first
    n_frames % world ranks get one extra frame."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], world: int, rank: int) -> Dict[str, torch.Tensor]:
    """Slice every [B,...] tensor of an AutoMoE batch dict to this rank's frames."""
    n = batch["image"].shape[0]
    s, c = shard_range(n, world, rank)
    return {k: (v[s:s + c] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def job_throughput(frames_local: int, ms_local: float, device=None) -> float:
    """Whole-job frames/s = frames of all ranks / slowest rank's device time (max over ranks)."""
    import torch.distributed as dist
    t = torch.tensor([ms_local, float(frames_local)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        tmax = t[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        fsum = t[1:].clone()
        dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        return fsum.item() / (tmax.item() / 1e3)
    return frames_local / (ms_local / 1e3)
