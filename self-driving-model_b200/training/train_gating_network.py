"""Gating / policy training step — drop-in for the hot part of training/train_gating_network.py
(compute_gating_losses :21-74, the body of train_one_epoch :93-105) on the sm_100a kernels.

Two ways to use it:

* the reference's trainer unchanged: `pred = model(batch)` (train mode, experts frozen) records the
  autograd graph over the kernels of training/functional.py; `compute_gating_losses` here returns the
  same dict from ONE fused kernel; `loss.backward()`, `clip_grad_norm_`, `torch.optim.AdamW` and
  `DistributedDataParallel` then work as they do on the reference;
* `FlatAdamW`: parameters, gradients and both Adam moments live in four flat fp32 buffers, so the
  data-parallel exchange is ONE NCCL all-reduce (SUM) of 11.5 MB over NVLink, and global-norm clip +
  AdamW is one reduction + one update kernel (amoe_sq_norm, amoe_fused_clip_adamw).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from .._cabi import check, ctx, lib
from .._ops import ptr, stream_ptr
from .functional import LOSS_NAMES, _GatingLoss, device_dropout_seed


def compute_gating_losses(pred: Dict[str, torch.Tensor], target_wp: torch.Tensor, target_spd: torch.Tensor,
                          config: Dict) -> Dict[str, torch.Tensor]:
    """Same signature, keys and values as the reference's compute_gating_losses."""
    wp = pred["waypoints"]
    pred_spd = pred.get("speed_seq", pred.get("speed"))
    # speed term selection as train_gating_network.py:28-37
    if pred_spd is not None and pred_spd.dim() == 2 and target_spd.dim() == 2 and pred_spd.size(1) == target_spd.size(1):
        mode, spd = 1, pred_spd
    else:
        pred_last = pred.get("speed")
        if pred_last is not None and pred_last.dim() == 2 and pred_last.size(1) == 1:
            mode, spd = 2, pred_last
        else:
            mode, spd = 0, None
    coef = [config.get('ade_weight', 1.0), config.get('fde_weight', 2.0), config.get('speed_weight', 0.2),
            config.get('smoothness_weight', 0.1), config.get('load_balancing_weight', 0.01),
            config.get('entropy_weight', 0.001)]
    tspd = target_spd if target_spd.dim() == 2 else target_spd.reshape(target_spd.size(0), -1)
    losses = _GatingLoss.apply(wp, spd, pred["expert_weights"], target_wp, tspd if mode else None, mode, coef,
                               bool(config.get('use_load_balancing', True)), bool(config.get('use_entropy_loss', True)))
    # total_loss is the differentiable output (what train_one_epoch back-propagates); the six terms are reported values
    # (detached: differentiating one of them alone is not supported by the fused kernel and must not silently give zeros)
    report = losses.detach()
    return {name: (losses[0] if i == 0 else report[i]) for i, name in enumerate(LOSS_NAMES)}


def allreduce_flat_(flat_grad: torch.Tensor, group=None) -> float:
    """SUM all-reduce of the flat gradient buffer across data-parallel ranks (NCCL over NVLink on GPUs, any
    torch.distributed backend otherwise).  Returns the scale (1/world) the optimizer applies."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world
    return 1.0


class FlatAdamW:
    """torch.optim.AdamW(params, lr, betas, eps, weight_decay) + clip_grad_norm_(max_norm) on flat buffers.

    The trainable parameters are re-pointed at slices of one flat fp32 buffer (their values are kept),
    `.grad` of every parameter is a slice of a second one, so autograd accumulates straight into the
    buffer that is all-reduced.  step() = all-reduce -> squared-norm reduction -> fused clip+AdamW.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_norm: Optional[float] = 1.0, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdamW got no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("automoe_b200 has no CPU path: FlatAdamW needs parameters on a CUDA (sm_100a) device")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm, self.group = lr, betas, eps, weight_decay, max_norm, group
        self.offsets, n = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatAdamW takes fp32 parameters on one device")
            self.offsets.append(n)
            n += (p.numel() + 3) & ~3          # 16-byte aligned slices
        self.n = n
        self.flat_param = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, off in zip(self.params, self.offsets):
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
        self._ws = torch.empty(4 * 148, device=dev, dtype=torch.float32)
        self._norm = torch.zeros(2, device=dev, dtype=torch.float32)
        self.step_count = 0
        self._step_dev: Optional[torch.Tensor] = None      # device-side step count / dropout key (GraphedTrainStep)
        self._seed_dev: Optional[torch.Tensor] = None

    def enable_device_counters(self) -> None:
        """Keep the step count (AdamW bias correction) and the per-step part of the dropout key in device memory, advanced by
        tick(): what a step replayed as a CUDA graph needs (host integers would be frozen into the graph)."""
        if self._step_dev is None:
            dev = self.flat_param.device
            self._step_dev = torch.tensor([self.step_count], device=dev, dtype=torch.int32)
            self._seed_dev = torch.zeros(1, device=dev, dtype=torch.int64)

    def tick(self) -> None:
        """Start of a step: advance the device-side counters (no-op without enable_device_counters)."""
        if self._step_dev is not None:
            dev = self.flat_param.device
            check(lib().amoe_train_tick(ctx(dev), ptr(self._step_dev), ptr(self._seed_dev), stream_ptr(dev)), "train_tick")

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):      # autograd may have replaced .grad (e.g. after set_to_none)
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + off * 4:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    def total_norm(self) -> torch.Tensor:
        """L2 norm of the SUMMED (all-reduced, not yet averaged) gradient of the last step(); the clip inside the fused
        kernel applies to norm / world, the averaged gradient's norm."""
        return self._norm[1]

    @torch.no_grad()
    def step(self):
        dev = self.flat_param.device
        for p, off in zip(self.params, self.offsets):
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + off * 4:
                self.flat_grad[off:off + p.numel()].view_as(p).copy_(p.grad)      # foreign .grad: fold it in
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
        scale = allreduce_flat_(self.flat_grad, self.group)
        h, st = ctx(dev), stream_ptr(dev)
        self.step_count += 1
        clip = self.max_norm is not None and self.max_norm > 0
        check(lib().amoe_sq_norm(h, ptr(self.flat_grad), self.n, ptr(self._ws), self._ws.numel(), ptr(self._norm), st), "sq_norm")
        if self._step_dev is not None:
            check(lib().amoe_fused_clip_adamw_dstep(h, ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg),
                                                    ptr(self.exp_avg_sq), self.n, ptr(self._norm), scale,
                                                    float(self.max_norm) if clip else 0.0, self.lr, self.betas[0], self.betas[1],
                                                    self.eps, self.weight_decay, ptr(self._step_dev), st), "fused_clip_adamw_dstep")
        else:
            check(lib().amoe_fused_clip_adamw(h, ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                              self.n, ptr(self._norm), scale, float(self.max_norm) if clip else 0.0, self.lr,
                                              self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count, st),
                  "fused_clip_adamw")
        # the kernel wrote through raw pointers: tell torch (cached weight packs key on Tensor._version)
        torch.autograd.graph.increment_version(self.params)


def broadcast_buffers_(model, group=None, src: int = 0) -> None:
    """What DistributedDataParallel(broadcast_buffers=True) does at the start of every forward: every rank takes rank 0's
    module buffers (BatchNorm running statistics and step counters), so replicas cannot drift apart.  One coalesced broadcast
    per dtype.  No-op outside torch.distributed / at world size 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    # Frozen experts on running statistics (frozen_experts_eval) never write their buffers: replicas cannot drift there, and
    # copying them would bump Tensor._version every step - the cached inference packs of the experts would be rebuilt per step
    skip = set()
    if getattr(model, "frozen_experts_eval", False) and hasattr(model, "experts"):
        skip = {id(b) for b in model.experts.buffers()}
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for b in model.buffers():
        if id(b) not in skip:
            by_dtype.setdefault(b.dtype, []).append(b)
    with torch.no_grad():
        for bufs in by_dtype.values():
            flat = torch.cat([b.reshape(-1) for b in bufs])
            dist.broadcast(flat, src, group=group)
            off = 0
            for b in bufs:
                b.copy_(flat[off:off + b.numel()].view_as(b))
                off += b.numel()


def freeze_for_gating_training(model) -> List[torch.nn.Parameter]:
    """model.freeze_experts() + the list of parameters the gating trainer optimises (2,870,657 for the
    3-expert configuration)."""
    model.freeze_experts()
    return [p for p in model.parameters() if p.requires_grad]


def train_step(model, batch: Dict[str, torch.Tensor], optimizer: FlatAdamW, config: Dict) -> Dict[str, torch.Tensor]:
    """One iteration of train_one_epoch (train_gating_network.py:93-105): zero_grad, forward, losses,
    backward, (all-reduce,) clip 1.0, AdamW.  Returns the loss dict (device tensors; no host sync)."""
    optimizer.zero_grad()
    optimizer.tick()
    broadcast_buffers_(model, getattr(optimizer, "group", None))      # DDP's per-forward buffer sync (no-op on one rank)
    with device_dropout_seed(optimizer._seed_dev):
        pred = model(batch)
    losses = compute_gating_losses(pred, batch["waypoints"], batch["speed"], config)
    losses["total_loss"].backward()
    optimizer.step()
    return {k: v.detach() for k, v in losses.items()}


class GraphedTrainStep:
    """train_step captured ONCE as a CUDA graph and replayed: forward in train mode, fused losses, backward,
    gradient all-reduce, clip + AdamW - about 540 launches per step whose enqueue costs the host as long as they run.

    What a replay cannot take from the host lives on the device: the optimizer's step count and the per-step part of the
    dropout key (FlatAdamW.enable_device_counters; amoe_train_tick advances both inside the graph), so bias correction and
    dropout masks differ from step to step exactly as in eager mode.  Construction runs `warmup` eager steps on `batch` (weight
    packs, allocator, NCCL) and then restores parameters, moments, module buffers and counters, so building the graph leaves
    the training state untouched.  Inputs are copied into static buffers per call; the returned loss dict aliases static
    device tensors (overwritten by the next call).  Shapes and the module's train/eval flags are fixed at capture time.
    Autograd graphs of earlier eager steps must be released before construction (see the error raised when capture fails).
    Frozen-expert weights enter the graph as packed at capture time: build a new GraphedTrainStep after loading other experts.
    """

    def __init__(self, model, batch: Dict[str, torch.Tensor], optimizer: FlatAdamW, config: Dict, warmup: int = 3,
                 autocast_dtype: Optional[torch.dtype] = None):
        from .._cabi import launch_count
        dev = optimizer.flat_param.device
        self.model, self.optimizer, self.config, self.autocast_dtype = model, optimizer, config, autocast_dtype
        self.static_in = {k: v.to(dev).clone() for k, v in batch.items() if torch.is_tensor(v)}
        optimizer.enable_device_counters()
        saved = [t.clone() for t in (optimizer.flat_param, optimizer.exp_avg, optimizer.exp_avg_sq, optimizer._step_dev,
                                     optimizer._seed_dev)]
        buffers = [(b, b.clone()) for b in model.buffers()]
        step_count = optimizer.step_count
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # staging caches keyed on Tensor._version must miss during capture, or the replay would keep the warm-up's frames
        torch.autograd.graph.increment_version(list(self.static_in.values()))
        self.graph = torch.cuda.CUDAGraph()
        n0 = launch_count(dev)
        try:
            with torch.cuda.graph(self.graph):
                self.losses = self._eager()
        except Exception as e:
            raise RuntimeError(
                "GraphedTrainStep: capturing the training step failed.  The usual cause is an autograd graph of an EARLIER eager "
                "step that is still alive (a kept loss / prediction tensor): its AccumulateGrad nodes stay bound to the stream "
                "they were created on, and the captured backward may not synchronise with the legacy default stream.  Drop those "
                "references (del loss, pred; gc.collect()) before building the graph.") from e
        self.launches_per_replay = launch_count(dev) - n0
        # The graph reads tensors that live outside its memory pool and are owned by caches: packed / split weights of the frozen
        # experts (their VALUES are baked into the graph as packed now: re-capture after loading other expert weights).  A later
        # eager forward may rebuild such a cache entry and drop the old tensors; holding them here keeps replays from reading
        # freed memory.
        self._keepalive = [dict(getattr(model, "_expert_packs", {}) or {})]
        self._keepalive += [dict(m._packs) for m in model.modules() if isinstance(getattr(m, "_packs", None), dict)]
        for p in model.parameters():
            for attr in ("_amoe_split6", "_amoe_split6_grouped", "_amoe_stem3"):
                c = getattr(p, attr, None)
                if c is not None:
                    self._keepalive.append(dict(c) if isinstance(c, dict) else c)
        with torch.no_grad():                       # nothing of the above counts as training
            for t, s in zip((optimizer.flat_param, optimizer.exp_avg, optimizer.exp_avg_sq, optimizer._step_dev,
                             optimizer._seed_dev), saved):
                t.copy_(s)
            for b, s in buffers:
                b.copy_(s)
            optimizer.flat_grad.zero_()
        optimizer.step_count = step_count
        torch.autograd.graph.increment_version(optimizer.params)

    def _eager(self):
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                return train_step(self.model, self.static_in, self.optimizer, self.config)
        return train_step(self.model, self.static_in, self.optimizer, self.config)

    def __call__(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        for k, dst in self.static_in.items():
            src = batch[k]
            if src.shape != dst.shape:
                raise ValueError(f"GraphedTrainStep was captured with {k}{tuple(dst.shape)}, got {tuple(src.shape)}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.optimizer.step_count += 1
        # the replay wrote parameters through raw pointers: cached weight packs key on Tensor._version
        torch.autograd.graph.increment_version(self.optimizer.params)
        return dict(self.losses)
