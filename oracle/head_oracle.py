"""TEST INFRASTRUCTURE — CPU restatement of ``TrajectoryHead.forward_test``.

A function-by-function restatement (plain torch CPU ops on a state dict, no
``nn.Module`` from the reference) of the planning-head inference path of
seulbinHwang/DiffusionDrive, *as written*: four dense 3x3 convolutions, the
full multi-head attentions, the per-scene time MLP.  It is the checker for the
CUDA path and the CPU baseline of bench.py; it is never the product path.

Pinning: the reference ships no test or golden vector for this path
(SURVEY.md §4), so the oracle is pinned against the *live* reference module,
imported in the build container by oracle/ref_import.py; the outputs are
committed under tests/golden/ (see oracle/make_golden.py) and
tests/test_oracle_golden.py re-checks this file against them.  The DDIM
arithmetic is third-party and version-unpinned upstream: see oracle/ddim.py
("parity unpinned" at that boundary).

Every function cites the reference lines it follows (paths relative to
/root/reference/navsim/agents/diffusiondrive/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .ddim import DDIMSchedulerRestated

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- helpers
def norm_odo(odo: torch.Tensor) -> torch.Tensor:
    """transfuser_model_v2.py:480-489 (heading slice empty on 2-channel input)."""
    x = odo[..., 0:1]
    y = odo[..., 1:2]
    h = odo[..., 2:3]
    x = 2 * (x + 1.2) / 56.9 - 1
    y = 2 * (y + 20) / 46 - 1
    h = 2 * (h + 2) / 3.9 - 1
    return torch.cat([x, y, h], dim=-1)


def denorm_odo(odo: torch.Tensor) -> torch.Tensor:
    """transfuser_model_v2.py:491-500."""
    x = odo[..., 0:1]
    y = odo[..., 1:2]
    h = odo[..., 2:3]
    x = (x + 1) / 2 * 56.9 - 1.2
    y = (y + 1) / 2 * 46 - 20
    h = (h + 1) / 2 * 3.9 - 2
    return torch.cat([x, y, h], dim=-1)


def gen_sineembed_for_position(pos: torch.Tensor, hidden_dim: int = 64) -> torch.Tensor:
    """modules/blocks.py:22-40 — DAB-DETR sine embedding; note the (y, x) output order
    and that ``dim_t`` is float32 regardless of the input dtype."""
    half = hidden_dim // 2
    scale = 2 * math.pi
    dim_t = torch.arange(half, dtype=torch.float32, device=pos.device)
    dim_t = 10000 ** (2 * (dim_t // 2) / half)
    x_embed = pos[..., 0] * scale
    y_embed = pos[..., 1] * scale
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=-1).flatten(-2)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=-1).flatten(-2)
    return torch.cat((pos_y, pos_x), dim=-1)


def sinusoidal_pos_emb(t: torch.Tensor, dim: int) -> torch.Tensor:
    """modules/conditional_unet1d.py:53-66."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=t.device) * -emb)
    emb = t[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], 1e-5)


def _mha(sd: SD, name: str, q: torch.Tensor, kv: torch.Tensor, heads: int) -> torch.Tensor:
    """nn.MultiheadAttention(batch_first=True, dropout 0) forward, output only
    (transfuser_model_v2.py:316-327; packed in_proj rows q|k|v)."""
    w = sd[name + ".in_proj_weight"]
    b = sd[name + ".in_proj_bias"]
    d = w.shape[1]
    hd = d // heads
    B, Lq, _ = q.shape
    Lk = kv.shape[1]
    qp = F.linear(q, w[:d], b[:d]).view(B, Lq, heads, hd).transpose(1, 2)
    kp = F.linear(kv, w[d:2 * d], b[d:2 * d]).view(B, Lk, heads, hd).transpose(1, 2)
    vp = F.linear(kv, w[2 * d:], b[2 * d:]).view(B, Lk, heads, hd).transpose(1, 2)
    att = torch.softmax((qp * (1.0 / math.sqrt(hd))) @ kp.transpose(-1, -2), dim=-1)
    o = (att @ vp).transpose(1, 2).reshape(B, Lq, d)
    return _lin(sd, name + ".out_proj", o)


# ----------------------------------------------------------------------------- blocks
def cross_bev_attention(sd: SD, pre: str, queries, traj_points, bev_feature,
                        lidar_max_x: float, lidar_max_y: float) -> torch.Tensor:
    """GridSampleCrossBEVAttention.forward, modules/blocks.py:88-129."""
    bs, nq, npts, _ = traj_points.shape
    nt = traj_points.clone()
    nt[..., 0] = nt[..., 0] / lidar_max_y            # :102-103
    nt[..., 1] = nt[..., 1] / lidar_max_x            # :104-105
    nt = nt[..., [1, 0]]                             # :107-108 swap -> (gx, gy)
    aw = _lin(sd, pre + "attention_weights", queries).view(bs, nq, npts).softmax(-1)   # :110-112
    value = F.relu(F.conv2d(bev_feature, sd[pre + "value_proj.0.weight"],
                            sd[pre + "value_proj.0.bias"], stride=1, padding=1))       # :114
    sampled = F.grid_sample(value, nt.view(bs, nq, npts, 2), mode="bilinear",
                            padding_mode="zeros", align_corners=False)                 # :117-122
    out = (aw.unsqueeze(1) * sampled).sum(dim=-1)     # :124-125
    out = out.permute(0, 2, 1).contiguous()           # :126
    out = _lin(sd, pre + "output_proj", out)          # :127
    return out + queries                              # :129 (dropout identity in eval)


def decoder_layer(sd: SD, pre: str, traj_feature, points, bev_feature, agents_query,
                  ego_query, time_embed, heads: int, lmx: float, lmy: float,
                  trace: Optional[dict] = None, tkey: str = ""
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """CustomTransformerDecoderLayer.forward, transfuser_model_v2.py:343-382.
    ``trace`` (tests only) records the intermediate activations under ``tkey``."""
    def rec(name, t):
        if trace is not None:
            trace[tkey + name] = t.detach().clone()
    f = cross_bev_attention(sd, pre + "cross_bev_attention.", traj_feature, points,
                            bev_feature, lmx, lmy)                                     # :353-354
    rec("x1", f)
    f = f + _mha(sd, pre + "cross_agent_attention", f, agents_query, heads)            # :355-357
    f = _ln(sd, pre + "norm1", f)                                                      # :358
    f = f + _mha(sd, pre + "cross_ego_attention", f, ego_query, heads)                 # :363-364
    f = _ln(sd, pre + "norm2", f)                                                      # :365
    rec("x2", f)
    f = _ln(sd, pre + "norm3",
            _lin(sd, pre + "ffn.2", F.relu(_lin(sd, pre + "ffn.0", f))))               # :368
    # ModulationLayer.forward :276-294
    ss = _lin(sd, pre + "time_modulation.scale_shift_mlp.1", F.mish(time_embed))
    scale, shift = ss.chunk(2, dim=-1)
    f = f * (1 + scale) + shift
    rec("x3", f)
    # DiffMotionPlanningRefinementModule.forward :244-256
    bs, modes, _ = f.shape
    c = f
    tp = pre + "task_decoder.plan_cls_branch."
    c = _ln(sd, tp + "2", F.relu(_lin(sd, tp + "0", c)))
    c = _ln(sd, tp + "5", F.relu(_lin(sd, tp + "3", c)))
    cls = _lin(sd, tp + "6", c).squeeze(-1)
    tr = pre + "task_decoder.plan_reg_branch."
    r = F.relu(_lin(sd, tr + "0", f))
    r = F.relu(_lin(sd, tr + "2", r))
    r = _lin(sd, tr + "4", r)
    num_poses = r.shape[-1] // 3
    reg = r.reshape(bs, modes, num_poses, 3).clone()
    reg[..., :2] = reg[..., :2] + points                                               # :378
    reg[..., 2] = reg[..., 2].tanh() * np.pi                                           # :379-380
    return reg, cls


def diff_decoder(sd: SD, num_layers: int, traj_feature, points, bev_feature, agents_query,
                 ego_query, time_embed, heads, lmx, lmy, trace=None, tkey=""):
    """CustomTransformerDecoder.forward, transfuser_model_v2.py:404-425: every layer sees
    the SAME traj_feature; only the points chain."""
    regs: List[torch.Tensor] = []
    clss: List[torch.Tensor] = []
    pts = points
    for l in range(num_layers):
        reg, cls = decoder_layer(sd, f"diff_decoder.layers.{l}.", traj_feature, pts, bev_feature,
                                 agents_query, ego_query, time_embed, heads, lmx, lmy,
                                 trace, f"{tkey}l{l}.")
        if trace is not None:
            trace[f"{tkey}l{l}.reg"] = reg.detach().clone()
            trace[f"{tkey}l{l}.cls"] = cls.detach().clone()
        regs.append(reg)
        clss.append(cls)
        pts = reg[..., :2].clone().detach()                                            # :424
    return regs, clss


# ----------------------------------------------------------------------------- the path
@torch.no_grad()
def forward_test(sd: SD, ego_query: torch.Tensor, agents_query: torch.Tensor,
                 bev_feature: torch.Tensor, noise: torch.Tensor, *, num_layers: int = 2,
                 step_num: int = 2, trunc_timestep: int = 8, heads: int = 8,
                 lidar_max_x: float = 32.0, lidar_max_y: float = 32.0,
                 dtype: Optional[torch.dtype] = None,
                 trace: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """TrajectoryHead.forward_test, transfuser_model_v2.py:578-641, with the noise of
    :593 injected.  Returns the reference's ``trajectory`` plus the locals
    ``poses_reg`` / ``poses_cls`` of the last denoise step (:630-631) as
    ``trajectory_modes`` / ``trajectory_scores`` and ``mode_idx``.

    ``dtype=torch.float64`` runs the same arithmetic in double (truth run); the sine
    tables stay float32-born as in the reference.
    """
    if dtype is not None:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        ego_query, agents_query, bev_feature, noise = (
            t.to(dtype) for t in (ego_query, agents_query, bev_feature, noise))
    work_dtype = ego_query.dtype
    sched = DDIMSchedulerRestated(num_train_timesteps=1000, beta_schedule="scaled_linear",
                                  prediction_type="sample")
    bs = ego_query.shape[0]
    sched.set_timesteps(1000)                                                          # :584
    step_ratio = 20 / step_num                                                         # :585
    roll = (np.arange(0, step_num) * step_ratio).round()[::-1].copy().astype(np.int64)  # :586-587
    plan_anchor = sd["plan_anchor"].unsqueeze(0).repeat(bs, 1, 1, 1)                   # :591
    img = norm_odo(plan_anchor)                                                        # :592
    trunc = torch.ones((bs,), dtype=torch.long) * trunc_timestep                       # :594
    img = sched.add_noise(original_samples=img, noise=noise, timesteps=trunc)          # :595-597
    modes = img.shape[1]
    d_model = sd["plan_anchor_encoder.3.weight"].shape[0]
    poses_reg = poses_cls = None
    for si, k in enumerate(roll):                                                      # :600
        x_boxes = torch.clamp(img, min=-1, max=1)                                      # :601
        pts = denorm_odo(x_boxes)                                                      # :602
        emb = gen_sineembed_for_position(pts, hidden_dim=64).flatten(-2)               # :605-607
        emb = emb.to(work_dtype)
        f = _lin(sd, "plan_anchor_encoder.0", emb)                                     # :608 (459-462)
        f = _ln(sd, "plan_anchor_encoder.2", F.relu(f))
        f = _lin(sd, "plan_anchor_encoder.3", f).view(bs, modes, -1)                   # :609
        ts = torch.tensor([int(k)], dtype=torch.long).expand(bs)                       # :611-621
        te = sinusoidal_pos_emb(ts, d_model).to(work_dtype)                            # :622 (463-468)
        te = _lin(sd, "time_mlp.3", F.mish(_lin(sd, "time_mlp.1", te))).view(bs, 1, -1)  # :623
        if trace is not None:
            trace[f"s{si}.img"] = img.clone()
            trace[f"s{si}.pts"] = pts.clone()
            trace[f"s{si}.q0"] = f.clone()
        regs, clss = diff_decoder(sd, num_layers, f, pts, bev_feature, agents_query,
                                  ego_query, te, heads, lidar_max_x, lidar_max_y,
                                  trace, f"s{si}.")                                    # :626-629
        poses_reg, poses_cls = regs[-1], clss[-1]                                      # :630-631
        x_start = norm_odo(poses_reg[..., :2])                                         # :632-633
        img = sched.step(model_output=x_start, timestep=k, sample=img).prev_sample     # :634-636
    mode_idx = poses_cls.argmax(dim=-1)                                                # :637
    gather_idx = mode_idx[..., None, None, None].repeat(1, 1, poses_reg.shape[2], 3)   # :638-639
    best = torch.gather(poses_reg, 1, gather_idx).squeeze(1)                           # :640
    return {"trajectory": best, "trajectory_modes": poses_reg, "trajectory_scores": poses_cls,
            "mode_idx": mode_idx}
