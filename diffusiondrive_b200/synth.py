"""Synthetic weights, anchors, backbone features and DDIM noise.

The real k-means anchor file and the trained checkpoint are release downloads
(docs/train_eval.md:13,30 of the reference) and unreachable offline, so every
test and benchmark of this repo runs on tensors regenerated from the seeds
below.  All draws use a CPU ``torch.Generator`` so that the GPU box reproduces
bit-identical inputs from the seed alone.

Weights are keyed by the reference's state-dict names
(transfuser_model_v2.py:455-476) and can be loaded unchanged into the live
reference ``TrajectoryHead`` (done by oracle/make_golden.py in the build
container) and into ``diffusiondrive_b200.TrajectoryHead``.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

SEED_WEIGHTS = 0
SEED_FEATURES = 1000
SEED_NOISE = 2000
SEED_THROUGHPUT = 3000


def make_anchors(num_anchors: int = 20, num_poses: int = 8) -> np.ndarray:
    """Constant-curvature arc anchors, (A, P, 2) float32 metres (x forward, y left).

    20 anchors: 5 speeds x 4 curvatures; 64 anchors: 8 speeds x 8 curvatures.
    About 10 % of the points lie beyond the 32 m BEV edge, which exercises the
    zero padding of ``grid_sample`` (modules/blocks.py:117-122).
    """
    if num_anchors == 20:
        speeds = [1.5, 4.5, 7.5, 10.5, 13.5]
        curvs = [0.0, 0.015, -0.015, 0.05]
    elif num_anchors == 64:
        speeds = [1.5, 3.0, 4.5, 6.0, 7.5, 9.0, 10.5, 13.5]
        curvs = [0.0, 0.008, -0.008, 0.015, -0.015, 0.03, -0.03, 0.05]
    else:
        n_s = int(math.ceil(math.sqrt(num_anchors)))
        speeds = list(np.linspace(1.5, 13.5, n_s))
        curvs = list(np.linspace(-0.05, 0.05, int(math.ceil(num_anchors / n_s))))
    t = 0.5 * (np.arange(num_poses, dtype=np.float64) + 1.0)
    out = []
    for v in speeds:
        for k in curvs:
            s = v * t
            if k == 0.0:
                x, y = s, np.zeros_like(s)
            else:
                th = k * s
                x, y = np.sin(th) / k, (1.0 - np.cos(th)) / k
            out.append(np.stack([x, y], -1))
    return np.asarray(out[:num_anchors], dtype=np.float32)


def _param_shapes(num_layers: int, d: int, f: int, p: int) -> Dict[str, Tuple[int, ...]]:
    """State-dict names and shapes of the reference head (minus ``plan_anchor``)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def ln(name):
        s[name + ".weight"] = (d,)
        s[name + ".bias"] = (d,)

    lin("plan_anchor_encoder.0", d, 512)
    ln("plan_anchor_encoder.2")
    lin("plan_anchor_encoder.3", d, d)
    lin("time_mlp.1", 4 * d, d)
    lin("time_mlp.3", d, 4 * d)
    for l in range(num_layers):
        L = f"diff_decoder.layers.{l}."
        lin(L + "cross_bev_attention.attention_weights", p, d)
        lin(L + "cross_bev_attention.output_proj", d, d)
        s[L + "cross_bev_attention.value_proj.0.weight"] = (256, 256, 3, 3)
        s[L + "cross_bev_attention.value_proj.0.bias"] = (256,)
        for att in ("cross_agent_attention", "cross_ego_attention"):
            s[L + att + ".in_proj_weight"] = (3 * d, d)
            s[L + att + ".in_proj_bias"] = (3 * d,)
            lin(L + att + ".out_proj", d, d)
        lin(L + "ffn.0", f, d)
        lin(L + "ffn.2", d, f)
        ln(L + "norm1")
        ln(L + "norm2")
        ln(L + "norm3")
        lin(L + "time_modulation.scale_shift_mlp.1", 2 * d, d)
        lin(L + "task_decoder.plan_cls_branch.0", d, d)
        ln(L + "task_decoder.plan_cls_branch.2")
        lin(L + "task_decoder.plan_cls_branch.3", d, d)
        ln(L + "task_decoder.plan_cls_branch.5")
        lin(L + "task_decoder.plan_cls_branch.6", 1, d)
        lin(L + "task_decoder.plan_reg_branch.0", d, d)
        lin(L + "task_decoder.plan_reg_branch.2", d, d)
        lin(L + "task_decoder.plan_reg_branch.4", p * 3, d)
    return s


def make_state_dict(seed: int = SEED_WEIGHTS, num_layers: int = 2, num_anchors: int = 20,
                    d_model: int = 256, d_ffn: int = 1024, num_poses: int = 8
                    ) -> Dict[str, torch.Tensor]:
    """Random weights under the reference's parameter names.

    Follows the spirit of the default inits (uniform +-1/sqrt(fan_in)) but
    defeats the two degenerate ones named in SURVEY.md §0.10: the zero-init
    ``attention_weights`` and the deep-copied (identical) decoder layers.  Every
    tensor is drawn independently; LayerNorm affine terms are perturbed.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {
        "plan_anchor": torch.from_numpy(make_anchors(num_anchors, num_poses))}
    for name, shape in _param_shapes(num_layers, d_model, d_ffn, num_poses).items():
        leaf = name.rsplit(".", 1)[-1]
        is_ln = (len(shape) == 1 and leaf == "weight") or any(
            name.endswith(t + "." + leaf) for t in (
                "norm1", "norm2", "norm3", "plan_anchor_encoder.2",
                "plan_cls_branch.2", "plan_cls_branch.5"))
        if is_ln:
            base = 1.0 if leaf == "weight" else 0.0
            t = base + 0.1 * torch.randn(shape, generator=g)
        elif "attention_weights" in name:
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("plan_cls_branch.6.bias"):
            # bias_init_with_prob(0.01), modules/blocks.py:43-46
            t = torch.full(shape, float(-np.log((1 - 0.01) / 0.01)))
        else:
            if len(shape) == 1:           # bias of a Linear / conv / packed in_proj
                fan_in = d_model if "ffn.2" not in name and "time_mlp.3" not in name else d_ffn
                if "value_proj" in name:
                    fan_in = 256 * 9
                if "plan_anchor_encoder.0" in name:
                    fan_in = 512
            else:
                fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[name] = t.to(torch.float32).contiguous()
    return sd


def make_features(batch: int, seed: int = SEED_FEATURES, num_agents: int = 30, d_model: int = 256,
                  bev_c: int = 256, bev_h: int = 64, bev_w: int = 64, start: int = 0
                  ) -> Dict[str, torch.Tensor]:
    """iid N(0,1) stand-ins for the three LayerNorm-ed backbone outputs the head reads.

    Shapes as produced by V2TransfuserModel.forward (transfuser_model_v2.py:136-156):
    ego_query (B,1,D), agents_query (B,Na,D), bev_feature (B,C,H,W) NCHW contiguous,
    status_encoding (B,1,D) (dead input, kept for the signature).
    Scenes are drawn one at a time so that scene ``i`` is the same tensor for
    every batch size (a B=1 run is the first scene of a B=256 run); ``start`` draws scenes
    ``start .. start + batch - 1`` of that sequence.
    """
    ego = torch.empty(batch, 1, d_model)
    agents = torch.empty(batch, num_agents, d_model)
    bev = torch.empty(batch, bev_c, bev_h, bev_w)
    status = torch.empty(batch, 1, d_model)
    for i in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 1_000_003 + start + i)
        ego[i] = torch.randn(1, d_model, generator=g)
        agents[i] = torch.randn(num_agents, d_model, generator=g)
        status[i] = torch.randn(1, d_model, generator=g)
        bev[i] = torch.randn(bev_c, bev_h, bev_w, generator=g)
    return {"ego_query": ego, "agents_query": agents, "bev_feature": bev,
            "status_encoding": status}


def make_noise(batch: int, seed: int = SEED_NOISE, num_anchors: int = 20, num_poses: int = 8,
               start: int = 0) -> torch.Tensor:
    """Host-generated DDIM noise (B, A, P, 2), injected in place of the reference's
    ``torch.randn(img.shape)`` (transfuser_model_v2.py:593)."""
    out = torch.empty(batch, num_anchors, num_poses, 2)
    for i in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 1_000_003 + start + i)
        out[i] = torch.randn(num_anchors, num_poses, 2, generator=g)
    return out


SEED_AGENT = 4000
AGENT_GAIN = 1.0          # variance gain of the random conv / linear weights (see make_agent_state_dict)


def make_agent_state_dict(agent, seed: int = SEED_AGENT) -> Dict[str, torch.Tensor]:
    """Random weights for every tensor of a ``DiffusionDriveAgent`` (same names as the reference
    ``V2TransfuserModel``), drawn in the module's own state_dict order from one seeded CPU
    generator; the planning head's tensors are those of ``make_state_dict`` so that the head
    fixtures stay comparable.  BatchNorm running statistics are perturbed (eval mode uses them)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    head_sd = make_state_dict()
    out: Dict[str, torch.Tensor] = {}
    for name, t in agent.state_dict().items():
        leaf = name.rsplit(".", 1)[-1]
        if name.startswith("_trajectory_head."):
            out[name] = head_sd[name[len("_trajectory_head."):]].clone()
        elif leaf == "num_batches_tracked":
            out[name] = torch.zeros_like(t)
        elif leaf == "running_var":
            out[name] = 1.0 + 0.2 * torch.rand(t.shape, generator=g)
        elif leaf == "running_mean":
            out[name] = 0.1 * torch.randn(t.shape, generator=g)
        elif t.dim() == 1 and leaf == "weight":            # BatchNorm / LayerNorm scale
            out[name] = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif "embedding" in name or leaf == "pos_emb":
            out[name] = 0.5 * torch.randn(t.shape, generator=g)
        elif t.dim() == 1:                                  # biases
            out[name] = 0.05 * torch.randn(t.shape, generator=g)
        else:
            fan_in = int(np.prod(t.shape[1:]))
            # He-style scale keeps activations O(1) through the ReLU stacks
            out[name] = torch.randn(t.shape, generator=g) * math.sqrt(AGENT_GAIN / fan_in)
        out[name] = out[name].to(t.dtype).contiguous()
    return out


def make_agent_inputs(batch: int, seed: int = SEED_AGENT + 1) -> Dict[str, torch.Tensor]:
    """camera (B,3,256,1024), LiDAR histogram (B,1,256,256), status (B,8): the feature dict of
    TransfuserFeatureBuilder (transfuser_features.py:25-139), iid stand-ins."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return {"camera_feature": torch.rand(batch, 3, 256, 1024, generator=g),
            "lidar_feature": (torch.rand(batch, 1, 256, 256, generator=g) > 0.9).float(),
            "status_feature": torch.randn(batch, 8, generator=g)}
