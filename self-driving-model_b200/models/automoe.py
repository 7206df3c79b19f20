"""AutoMoE composite model — drop-in for models/automoe.py (create_automoe_model,
AutoMoE.forward / get_expert_weights / load_expert_checkpoints / freeze_experts /
unfreeze_experts) with the same constructor config, attribute names, forward dict and
state_dict keys; the forward runs on hand-written sm_100a kernels through the C-ABI.

Forward schedule (reference automoe.py:189-233, re-designed):
  image NCHW fp32 -> NHWC (one kernel, shared by experts and policy)
  3 experts: grouped ResNet-18 trunks + heads (tcgen05 implicit-GEMM convs, BN folded),
             1x1 classifier + pooled mean, bilinear x32 writer for the full-res logits
  ONE fused kernel: context extractor + expert extractors + gating + softmax + combine
  policy: 4 convs + fused pool/fc/MLP-heads kernel
"""
from __future__ import annotations

import warnings
from typing import Dict, List

import torch
import torch.nn as nn

from .. import _ops
from ._gatepack import pack_gate_params, require_eval
from ._precision import resolve_dtype
from .context.context_features import create_context_extractor
from .experts import BDDDetectionExpert, BDDDrivableExpert, BDDSegmentationExpert, NuScenesExpert
from .experts._base import BDDExpertBase, get_trunk_pack, run_experts, side_stream
from .experts._trunk import (chunked_stem_layer1_supported, grouped_train_supported, params_stamp, run_stem_layer1_chunked,
                             run_trunk_train, run_trunks_train_grouped, stage_image, trunk_pool_pad)
from .experts.expert_extractors import create_expert_extractors
from .gating.gating_network import GatingNetwork
from .policy.trajectory_head import TrajectoryPolicy


class AutoMoE(nn.Module):
    """Complete AutoMoE: Mixture of Experts Self-Driving Model"""

    def __init__(self, expert_configs: List[Dict], gating_config: Dict, context_config: Dict,
                 policy_config: Dict, device: str = 'cuda', precision: str = 'auto'):
        super().__init__()
        self.device = device
        self.expert_configs = expert_configs
        self.gating_config = gating_config
        self.context_config = context_config
        self.policy_config = policy_config
        self.precision = precision

        self.experts = self._create_experts()
        self.expert_extractors = create_expert_extractors(expert_configs)
        self.context_extractor = create_context_extractor(context_config)
        self.gating_network = self._create_gating_network()
        self.policy_head = self._create_policy_head()
        self._expert_packs = {}
        self._gate_flat = {}
        self.to(device)

    def _create_experts(self) -> nn.ModuleList:
        experts = nn.ModuleList()
        for config in self.expert_configs:
            expert_type = config['type']
            if expert_type == 'detection':
                expert = BDDDetectionExpert(num_classes=config.get('num_classes', 10),
                                            pretrained_backbone=config.get('pretrained_backbone', True))
            elif expert_type == 'segmentation':
                expert = BDDSegmentationExpert(num_classes=config.get('num_classes', 19),
                                               pretrained_backbone=config.get('pretrained_backbone', True))
            elif expert_type == 'drivable':
                expert = BDDDrivableExpert(num_classes=config.get('num_classes', 3),
                                           pretrained_backbone=config.get('pretrained_backbone', True))
            elif expert_type == 'nuscenes':
                expert = NuScenesExpert(num_queries=config.get('num_queries', 100), fusion=config.get('fusion', 'concat'),
                                        use_lidar=config.get('use_lidar', False), use_tnet=config.get('use_tnet', False),
                                        bbox_dim=config.get('bbox_dim', 7),
                                        pretrained_backbone=config.get('pretrained_backbone', True))
            else:
                raise ValueError(f"Unknown expert type: {expert_type}")
            experts.append(expert)
        return experts

    def _create_gating_network(self) -> GatingNetwork:
        num_experts = len(self.expert_configs)
        expert_output_dims = [config.get('output_dim', 256) for config in self.expert_configs]
        # top_k / noise_* / apply_topk_at_eval in the JSON are not forwarded (automoe.py:83-91)
        return GatingNetwork(
            num_experts=num_experts,
            context_dim=self.context_config.get('context_dim', 64),
            expert_output_dims=expert_output_dims,
            processed_dim=self.gating_config.get('processed_dim', 256),
            hidden_dim=self.gating_config.get('hidden_dim', 128),
            temperature=self.gating_config.get('temperature', 1.0),
            use_softmax=self.gating_config.get('use_softmax', True))

    def _create_policy_head(self) -> TrajectoryPolicy:
        return TrajectoryPolicy(
            horizon=self.policy_config.get('num_waypoints', 10),
            context_dim=self.gating_config.get('processed_dim', 256),
            backbone_dim=self.policy_config.get('backbone_dim', 512))

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _last_col(t: torch.Tensor) -> torch.Tensor:
        """[B,1] passes through; [B,H] -> last column; higher rank -> flattened last element
        (automoe.py:108-134)."""
        if t.dim() == 2 and t.size(1) > 1:
            return t[:, -1:]
        if t.dim() > 2:
            return t.reshape(t.size(0), -1)[:, -1:]
        return t

    def _vehicle_state(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """[B,4] (speed, steering, throttle, brake) exactly as _extract_context_features builds it."""
        if self.context_config.get('type', 'simple') != 'simple':
            raise NotImplementedError("context type 'full' is dead code in the reference")
        speed = batch['speed']
        speed_in = speed[:, -1:] if (speed.dim() == 2 and speed.size(1) > 1) else speed
        if all(k in batch for k in ('speed', 'steering', 'throttle', 'brake')):
            cols = [speed_in] + [self._last_col(batch[k]) for k in ('steering', 'throttle', 'brake')]
        else:
            z = torch.zeros(speed_in.size(0), 1, device=speed_in.device)
            cols = [speed_in, z, z, z]
        return torch.cat([c.reshape(c.size(0), 1).float() for c in cols], dim=-1).contiguous()

    def _gate_params(self, device, n_ch, bf16_copy=False):
        """Flat fp32 parameter buffer of the fused gate kernel (cached; re-packed when a parameter changes).
        bf16_copy=True returns (flat, flat.bfloat16()) for the tensor-core variant of bf16 inference."""
        stamp = params_stamp([self.context_extractor, self.expert_extractors, self.gating_network])
        key = device.index
        g = self._gate_flat.get(key)
        if g is None or g[0] != stamp:
            flat = pack_gate_params(self.context_extractor, list(self.expert_extractors.extractors),
                                    self.gating_network, n_ch, self.context_extractor.context_dim,
                                    self.gating_network.hidden_dim, device)
            g = (stamp, flat, flat.to(torch.bfloat16))
            self._gate_flat[key] = g
        return (g[1], g[2]) if bf16_copy else g[1]

    def _fused_stem(self, device):
        """PackedStem of [expert stems..., policy conv1] (cached; re-packed when any of them changes)."""
        bdd = self._bdd_experts()
        mods = [e.backbone[0] for e in bdd] + [e.backbone[1] for e in bdd] + \
               [self.policy_head.backbone.net[0], self.policy_head.backbone.net[1]]
        stamp = params_stamp(mods)
        c = self._gate_flat.get(("stem", device.index))
        if c is None or c[0] != stamp:
            convs = [e.backbone[0] for e in bdd] + [self.policy_head.backbone.net[0]]
            bns = [e.backbone[1] for e in bdd] + [self.policy_head.backbone.net[1]]
            c = (stamp, _ops.pack_stem(convs, bns, device, relu=True))
            self._gate_flat[("stem", device.index)] = c
        return c[1]

    input_mean = _ops.IMAGENET_MEAN      # T.Normalize constants of inference/run_automoe.py:30
    input_std = _ops.IMAGENET_STD

    def _stage_u8(self, frames: torch.Tensor, dtype: torch.dtype):
        """uint8 HWC frames -> (image, x_nhwc).  bf16 tensor-core stem: one kernel writes the normalised, padded NHWC4
        frame directly and `image` is only a [B,3,H,W] shape carrier (a permuted view of the bytes, never read);
        otherwise the fp32 NCHW tensor of the reference transform is materialised and staged as usual."""
        if frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError(f"uint8 frames must be [B,H,W,3] (HWC RGB), got {tuple(frames.shape)}")
        H, W = frames.shape[1], frames.shape[2]
        if _ops.stem_mode(dtype, H, W) == "tc":
            return frames.permute(0, 3, 1, 2), _ops.stage_u8_stem(frames, self.input_mean, self.input_std)
        image = _ops.normalize_u8_nchw(frames, self.input_mean, self.input_std)
        return image, stage_image(image, dtype)

    def _bdd_experts(self):
        """The experts that share the grouped ResNet-18 + 2-conv-head launches (detection / segmentation / drivable)."""
        return [e for e in self.experts if isinstance(e, BDDExpertBase)]

    def _other_expert_features(self, image, dtype, x_nhwc, expert_outputs, n_ch_bdd):
        """Experts outside the grouped launches (nuScenes): run them, return (expert_outputs in config order, n_ch per expert
        with 0 for externally extracted features, ext_features [E,B,256] or None)."""
        E = len(self.experts)
        if len(expert_outputs) == E:
            return expert_outputs, n_ch_bdd, None
        outs, n_ch, bdd_i = [], [], 0
        ext = torch.zeros((E, image.shape[0], 256), device=image.device, dtype=torch.float32)
        for i, e in enumerate(self.experts):
            if isinstance(e, BDDExpertBase):
                outs.append(expert_outputs[bdd_i])
                n_ch.append(n_ch_bdd[bdd_i])
                bdd_i += 1
            else:
                out = e({'image': image}, _x_nhwc=x_nhwc, _dtype=dtype)
                with torch.no_grad():
                    ext[i] = self.expert_extractors.extractors[i](out)
                outs.append({k: v for k, v in out.items() if not k.startswith('_')})
                n_ch.append(0)
        return outs, n_ch, ext

    def _extract_context_features(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        state = self._vehicle_state(batch)
        return self.context_extractor(state[:, 0:1], state[:, 1:2], state[:, 2:3], state[:, 3:4])

    def _run_experts(self, batch: Dict[str, torch.Tensor]) -> List:
        """All experts on batch['image'] in grouped launches.  Unlike the reference
        (automoe.py:181-185) an expert failure raises instead of being masked by zeros."""
        dtype = resolve_dtype(self.precision)
        outs, aux = run_experts(self._bdd_experts(), batch['image'], dtype, self._expert_packs)
        return self._other_expert_features(batch['image'], dtype, None, outs, aux['n_ch'])[0]

    # ------------------------------------------------------------------ forward
    def _trainable_part(self):
        return [self.context_extractor, self.expert_extractors, self.gating_network, self.policy_head]

    # Reference train-mode semantics (training/train_gating_network.py:85 calls model.train() on the whole model):
    # the frozen experts' BatchNorm layers then normalise with BATCH statistics and keep updating their running
    # statistics, and every Dropout is active.  That is the default here too.  frozen_experts_eval = True is the
    # explicit opt-out: frozen experts keep their running statistics and run through the (much faster) inference
    # kernels - the numbers the reference gives after `model.experts.eval()`.
    frozen_experts_eval = False
    _warned_eval_grad = False

    def _run_experts_train_mode(self, image: torch.Tensor):
        """Frozen experts exactly as the reference runs them inside model.train(): batch-statistics BatchNorm with
        running-stat updates (fp32 training kernels, no autograd graph).  Returns (expert_outputs, pooled, n_ch)."""
        outs, pooled = [], []
        H, W = image.shape[2], image.shape[3]
        bdd = self._bdd_experts()
        with torch.no_grad():
            # same-geometry experts advance layer by layer in grouped launches (one conv / BatchNorm pass for all of them)
            lows = run_trunks_train_grouped(bdd, image) if grouped_train_supported(bdd, image) else None
            for gi, e in enumerate(bdd):
                low = lows[gi] if lows is not None else run_trunk_train(e, image)      # [B,h,w,N] NHWC fp32
                out = e.format_output_train(low, H, W)
                outs.append({k: v for k, v in out.items() if not k.startswith('_')} if isinstance(out, dict) else out)
                h, w = low.shape[1], low.shape[2]
                if e.upsample_to_input and (H % h != 0 or W % w != 0):
                    pooled.append(_ops.mean_hw_nchw(out))                # non-integer scale: pool the full-resolution map
                else:
                    pooled.append(low.mean(dim=(1, 2)))                  # == mean of the x32 bilinear map (integer scale)
        n_ch = [int(t.shape[1]) for t in pooled]
        return outs, torch.cat(pooled, dim=1).contiguous(), n_ch

    def _forward_train(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Forward that records an autograd graph over the gating / policy parameters
        (training/train_gating_network.py:96-100).  Experts must be frozen (freeze_experts()).  Experts in train
        mode (what model.train() gives) follow the reference: batch-statistics BatchNorm + running-stat updates;
        experts in eval mode, or frozen_experts_eval = True, run through the inference kernels with running statistics."""
        from ._train_forward import context_extractor_forward, extractor_forward, gating_forward, policy_forward
        if any(p.requires_grad for p in self.experts.parameters()):
            raise NotImplementedError("joint training of the experts (unfreeze_experts) is not implemented: "
                                      "call freeze_experts() - only the gating/policy part back-propagates")
        image = batch['image']
        if not image.is_cuda:
            raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
        if image.dtype == torch.uint8:
            image = _ops.normalize_u8_nchw(image, self.input_mean, self.input_std)
        dtype = resolve_dtype(self.precision)
        state = self._vehicle_state(batch).to(image.device)
        ref_mode = any(e.training for e in self.experts) and not self.frozen_experts_eval
        if ref_mode:
            expert_outputs, pooled, n_ch = self._run_experts_train_mode(image)
        else:
            with torch.no_grad():
                expert_outputs, aux = run_experts(self._bdd_experts(), image, dtype, self._expert_packs, frozen_eval=True)
            pooled, n_ch = aux['pooled'], aux['n_ch']
        ctx = context_extractor_forward(self.context_extractor, state)
        feats, off, bdd_i, outs_all = [], 0, 0, []
        for e, ext in zip(self.experts, self.expert_extractors.extractors):
            if isinstance(e, BDDExpertBase):
                n = n_ch[bdd_i]
                feats.append(extractor_forward(ext, pooled[:, off:off + n].contiguous()))
                outs_all.append(expert_outputs[bdd_i])
                off += n
                bdd_i += 1
            else:
                # nuScenes expert: frozen; in train mode it follows the reference (batch-statistics BatchNorm, active Dropout)
                was = e.training
                if not ref_mode:
                    e.eval()
                try:
                    out = e({'image': image}, _dtype=dtype)
                finally:
                    e.train(was)
                feats.append(ext(out))                    # differentiable extractor (training/functional.py shims)
                outs_all.append({k: v for k, v in out.items() if not k.startswith('_')})
        expert_outputs = outs_all
        g = gating_forward(self.gating_network, feats, ctx)
        pol = policy_forward(self.policy_head, image, g['combined_output'])
        speed_seq = pol['speed']
        return {
            'waypoints': pol['waypoints'],
            'speed': speed_seq[:, -1:].contiguous(),
            'speed_seq': speed_seq,
            'expert_weights': g['expert_weights'],
            'expert_outputs': expert_outputs,
            'context_features': ctx,
            'combined_features': g['combined_output'],
            'gate_logits': g['gate_logits'],
        }

    def forward(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        from ._train_forward import wants_grad
        if self.training:
            return self._forward_train(batch)
        if torch.is_grad_enabled() and any(wants_grad(m) for m in self._trainable_part()) and \
                not any(p.requires_grad for p in self.experts.parameters()):
            if not AutoMoE._warned_eval_grad:
                AutoMoE._warned_eval_grad = True
                warnings.warn("automoe_b200: eval-mode forward with autograd recording takes the differentiable fp32 path "
                              "(training kernels); wrap inference in torch.no_grad() for the tensor-core inference kernels")
            return self._forward_train(batch)      # eval semantics, differentiable (frozen experts)
        image = batch['image']
        if not image.is_cuda:
            raise RuntimeError("automoe_b200 has no CPU path: move the model and the batch to a CUDA (sm_100a) device")
        dtype = resolve_dtype(self.precision)
        state = self._vehicle_state(batch).to(image.device)
        if image.dtype == torch.uint8:
            # camera bytes [B,H,W,3]: ToTensor + Normalize (+ the stem's layout) happen on the device
            # (inference/run_automoe.py:25-31,41 does them per frame on the CPU and uploads fp32)
            image, x_nhwc = self._stage_u8(image, dtype)
        else:
            x_nhwc = stage_image(image, dtype)
        stem_out = pol1 = pooled = layer1 = None
        if _ops.stem_mode(dtype, image.shape[2], image.shape[3]) == "tc":
            # experts' stems + policy conv1 read the same frame: one GEMM with N = 3*64 + 32
            # (+ the experts' max-pool fused behind it when the geometry allows)
            fs = self._fused_stem(image.device)
            Bn, Hn, Wn = image.shape[0], image.shape[2], image.shape[3]
            bdd = self._bdd_experts()
            tp = get_trunk_pack(bdd, dtype, image.device, self._expert_packs)
            if chunked_stem_layer1_supported(tp, Bn, Hn, Wn):
                # stem+pool and layer1 walk the batch in L2-resident chunks (policy conv1 rides along)
                pol1 = torch.empty((Bn, Hn // 2, Wn // 2, fs.couts[-1]), device=image.device, dtype=torch.bfloat16)
                layer1 = run_stem_layer1_chunked(tp, fs, x_nhwc, Bn, Hn, Wn, rest_out=[pol1])
            elif _ops.stem_pool_supported(Hn, Wn):
                pooled, rest = _ops.stem_pool_forward(fs, x_nhwc, Bn, Hn, Wn, len(bdd),
                                                      trunk_pool_pad(tp, Hn, Wn))
                pol1 = rest[0]
            else:
                stem_out, pol1 = _ops.stem_forward(fs, x_nhwc, Bn, Hn, Wn, groups=[len(bdd), 1])

        # The policy backbone (conv2-4 + pooling) depends on the stem only: it is forked onto a second stream right behind the
        # last tensor-bound expert launch, beside the 1x1 heads and the gate kernel (latency-bound, 64 CTAs), and the policy
        # heads join it behind the gate.
        pol = {}

        def fork_policy_backbone():
            main = torch.cuda.current_stream(image.device)
            ps = side_stream(image.device, 1)
            ps.wait_stream(main)
            with torch.cuda.stream(ps):
                pol['feat'] = self.policy_head.backbone_features(image, _dtype=dtype, _conv1=pol1)
            pol1.record_stream(ps)
            pol['feat'].record_stream(main)
            pol['stream'] = ps

        expert_outputs, aux = run_experts(self._bdd_experts(), image, dtype, self._expert_packs, x_nhwc=x_nhwc,
                                          stem_out=stem_out, stem_pooled=pooled, layer1_out=layer1,
                                          overlap_outputs=_ops.overlap_outputs(),
                                          after_head3=fork_policy_backbone if (pol1 is not None and _ops.overlap_tail()) else None)
        expert_outputs, n_ch, ext = self._other_expert_features(image, dtype, x_nhwc, expert_outputs, aux['n_ch'])

        gn = self.gating_network
        gflat, gflat16 = self._gate_params(image.device, n_ch, bf16_copy=True)
        g = _ops.gate(state, aux['pooled'], gflat, n_ch,
                      self.context_extractor.context_dim, gn.hidden_dim, gn.temperature, mode=gn._gate_kind(),
                      params_bf16=gflat16 if _ops.mlp_tc(dtype) else None, ext_features=ext)

        if pol.get('stream') is not None:
            torch.cuda.current_stream(image.device).wait_stream(pol['stream'])
        policy_output = self.policy_head(image, context=g['combined'], _x_nhwc=x_nhwc, _dtype=dtype, _conv1=pol1,
                                         _feat=pol.get('feat'))
        if aux.get('join') is not None:      # full-resolution logits were written on the side stream
            torch.cuda.current_stream(image.device).wait_stream(aux['join'])
        speed_seq = policy_output.get('speed')
        speed_out = None
        if speed_seq is not None and speed_seq.dim() == 2:
            speed_out = speed_seq[:, -1:].contiguous()

        return {
            'waypoints': policy_output['waypoints'],
            'speed': speed_out if speed_out is not None else speed_seq,
            'speed_seq': speed_seq,
            'expert_weights': g['weights'],
            'expert_outputs': expert_outputs,
            'context_features': g['context'],
            'combined_features': g['combined'],
            'gate_logits': g['gate_logits'],
        }

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, batch: Dict[str, torch.Tensor], autocast_dtype=torch.bfloat16,
                clone_inputs: bool = True) -> "GraphedForward":
        """Capture forward(batch) into a CUDA graph on static copies of the inputs (fixed shapes).

        The forward is ~35 launches at batch 256 and ~70 when stem+layer1 walk the batch in L2-sized
        chunks; replaying a graph removes the per-launch host cost (ctypes + allocator), so short
        kernels queue back to back.  Weights are baked in as packed at capture time: re-capture after
        an optimizer step / load_state_dict.  clone_inputs=False captures on the given tensors
        themselves: the caller refills those buffers in place and calls the result without arguments."""
        return GraphedForward(self, batch, autocast_dtype, clone_inputs)

    def get_expert_weights(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Get expert weights without running experts (for analysis)"""
        context_features = self._extract_context_features(batch)
        return self.gating_network.get_expert_weights(context_features)

    def load_expert_checkpoints(self, checkpoint_paths: List[str]):
        """Load pre-trained expert checkpoints (automoe.py:240-267)"""
        if len(checkpoint_paths) != len(self.experts):
            raise ValueError(f"Expected {len(self.experts)} checkpoint paths, got {len(checkpoint_paths)}")
        for i, (expert, checkpoint_path) in enumerate(zip(self.experts, checkpoint_paths)):
            if checkpoint_path and checkpoint_path != '':
                try:
                    checkpoint = torch.load(checkpoint_path, map_location=self.device)
                    state_dict = checkpoint.get('model_state_dict', checkpoint)
                    expert.load_state_dict(state_dict)
                    print(f"Loaded checkpoint for expert {i}: {checkpoint_path}")
                except Exception as e:
                    warnings.warn(f"Failed to load checkpoint for expert {i}: {str(e)}")

    def freeze_experts(self):
        """Freeze expert parameters during gating network training"""
        for expert in self.experts:
            for param in expert.parameters():
                param.requires_grad = False

    def unfreeze_experts(self):
        """Unfreeze expert parameters for joint training"""
        for expert in self.experts:
            for param in expert.parameters():
                param.requires_grad = True


class GraphedForward:
    """Replayable CUDA graph of AutoMoE.forward on fixed input shapes (see AutoMoE.capture)."""

    def __init__(self, model: AutoMoE, batch: Dict[str, torch.Tensor], autocast_dtype=torch.bfloat16,
                 clone_inputs: bool = True):
        from .. import _cabi
        dev = batch['image'].device
        self.static_in = {k: (v.clone() if clone_inputs else v) for k, v in batch.items() if torch.is_tensor(v)}
        self.autocast_dtype = autocast_dtype
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(warm), torch.no_grad():      # packs + allocator warm before capture
            for _ in range(2):
                self._run(model)
        torch.cuda.current_stream(dev).wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _cabi.launch_count(dev)
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self._run(model)
        self.launches_per_replay = _cabi.launch_count(dev) - n0

    def _run(self, model):
        if self.autocast_dtype is None:
            return model(self.static_in)
        with torch.autocast("cuda", dtype=self.autocast_dtype):
            return model(self.static_in)

    def __call__(self, batch: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Copy `batch` into the static inputs (if given), replay, return the static outputs."""
        if batch is not None:
            for k, v in self.static_in.items():
                v.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.static_out


def create_automoe_model(config: Dict, device: str = 'cuda') -> AutoMoE:
    """Create AutoMoE model from configuration (same schema as models/configs/automoe/model_config.json;
    optional extra key "precision": "auto" | "bf16" | "fp32")."""
    return AutoMoE(
        expert_configs=config['experts'],
        gating_config=config['gating'],
        context_config=config['context'],
        policy_config=config['policy'],
        device=device,
        precision=config.get('precision', 'auto'),
    )
