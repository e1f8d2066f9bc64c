"""Training path of the planning head: ``forward_train``
(navsim/agents/diffusiondrive/transfuser_model_v2.py:520-576) and ``LossComputer``
(modules/multimodal_loss.py:119-168) as plain differentiable PyTorch on the parameters the drop-in
``TrajectoryHead`` holds.

This is NOT the accelerated path and not a fallback of it: inference (``forward_test``) only ever
runs behind the C ABI.  Training needs autograd through the decoder, which the CUDA engines do not
provide, so ``TrajectoryHead.forward`` in training mode delegates here, on whatever device the
module lives on -- the behaviour SURVEY.md section 8(b) asks of the boundary ("training mode must
keep delegating to reference-equivalent PyTorch").
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .config import ODO_X_OFF, ODO_X_RANGE, ODO_Y_OFF, ODO_Y_RANGE, SINE_HIDDEN


def norm_odo(x: torch.Tensor) -> torch.Tensor:
    """(..., 2) metres -> [-1, 1]  (:480-489; the heading slice is empty for 2-channel input)."""
    return torch.stack((2 * (x[..., 0] + ODO_X_OFF) / ODO_X_RANGE - 1,
                        2 * (x[..., 1] + ODO_Y_OFF) / ODO_Y_RANGE - 1), -1)


def denorm_odo(x: torch.Tensor) -> torch.Tensor:
    """Inverse of norm_odo (:491-500)."""
    return torch.stack(((x[..., 0] + 1) / 2 * ODO_X_RANGE - ODO_X_OFF,
                        (x[..., 1] + 1) / 2 * ODO_Y_RANGE - ODO_Y_OFF), -1)


def sine_embed(pts: torch.Tensor, hidden: int = SINE_HIDDEN) -> torch.Tensor:
    """gen_sineembed_for_position (modules/blocks.py:22-40): (..., 2) -> (..., hidden), y first."""
    half = hidden // 2
    i = torch.arange(half, dtype=torch.float32, device=pts.device)
    dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="floor") / half)

    def emb(v):
        e = (v * (2 * math.pi))[..., None] / dim_t
        return torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), -1).flatten(-2)
    return torch.cat((emb(pts[..., 1]), emb(pts[..., 0])), -1)


def time_embed(head, t: torch.Tensor) -> torch.Tensor:
    """SinusoidalPosEmb(256) + time_mlp (modules/conditional_unet1d.py:53-66, :463-468)."""
    d = head._d_model
    half = d // 2
    f = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    e = t[:, None] * f[None, :]
    return head.time_mlp(torch.cat((e.sin(), e.cos()), -1))


DROPOUT_P = 0.1    # the three nn.Dropout(0.1) of the training graph (modules/blocks.py:66, :307-308)


def bev_attention(p, queries, pts, bev, lidar_max_x: float, lidar_max_y: float, drop: float = 0.0):
    """GridSampleCrossBEVAttention.forward (modules/blocks.py:88-129)."""
    bs, nq, npts, _ = pts.shape
    grid = torch.stack((pts[..., 1] / lidar_max_x, pts[..., 0] / lidar_max_y), -1)
    w = p.attention_weights(queries).view(bs, nq, npts).softmax(-1)
    sampled = F.grid_sample(p.value_proj(bev), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    out = (w.unsqueeze(1) * sampled).sum(-1).permute(0, 2, 1).contiguous()
    return F.dropout(p.output_proj(out), drop, training=drop > 0) + queries


def decoder_layer(layer, feat, pts, bev, agents, ego, t_emb, lx, ly, drop: float = 0.0):
    """CustomTransformerDecoderLayer.forward (:343-382)."""
    x = bev_attention(layer.cross_bev_attention, feat, pts, bev, lx, ly, drop)
    x = layer.norm1(x + F.dropout(layer.cross_agent_attention(x, agents, agents)[0], drop, training=drop > 0))
    x = layer.norm2(x + F.dropout(layer.cross_ego_attention(x, ego, ego)[0], drop, training=drop > 0))
    x = layer.norm3(layer.ffn(x))
    scale, shift = layer.time_modulation.scale_shift_mlp(t_emb).chunk(2, -1)
    x = x * (1 + scale) + shift
    bs, modes = x.shape[:2]
    td = layer.task_decoder
    reg = td.plan_reg_branch(x).view(bs, modes, -1, 3)
    cls = td.plan_cls_branch(x).squeeze(-1)
    xy = reg[..., :2] + pts
    return torch.cat((xy, reg[..., 2:3].tanh() * math.pi), -1), cls


def decoder(head, feat, pts, bev, agents, ego, t_emb, drop: float = 0.0):
    """CustomTransformerDecoder.forward (:404-425): same traj_feature for every layer, the points
    chain through the (detached) regression output."""
    regs, clss = [], []
    for layer in head.diff_decoder.layers:
        reg, cls = decoder_layer(layer, feat, pts, bev, agents, ego, t_emb, head._lidar_max_x, head._lidar_max_y, drop)
        regs.append(reg)
        clss.append(cls)
        pts = reg[..., :2].clone().detach()
    return regs, clss


def sigmoid_focal_loss(pred, target, gamma: float = 2.0, alpha: float = 0.25):
    """py_sigmoid_focal_loss with reduction "mean" (modules/multimodal_loss.py:70-117)."""
    ps = pred.sigmoid()
    pt = (1 - ps) * target + ps * (1 - target)
    fw = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    return (F.binary_cross_entropy_with_logits(pred, target, reduction="none") * fw).mean()


def trajectory_loss(reg, cls, target_traj, plan_anchor, cls_weight: float = 10.0, reg_weight: float = 8.0):
    """LossComputer.forward (modules/multimodal_loss.py:128-168): focal loss towards the anchor
    nearest the target, L1 on that anchor's poses; weights transfuser_config.py:84-85."""
    bs, modes, ts, d = reg.shape
    dist = torch.linalg.norm(target_traj.unsqueeze(1)[..., :2] - plan_anchor, dim=-1).mean(-1)
    nearest = dist.argmin(-1)
    best = torch.gather(reg, 1, nearest[:, None, None, None].expand(-1, 1, ts, d)).squeeze(1)
    onehot = torch.zeros_like(cls).scatter_(1, nearest[:, None], 1.0)
    return cls_weight * sigmoid_focal_loss(cls, onehot) + reg_weight * F.l1_loss(best, target_traj)


def forward_train(head, ego_query, agents_query, bev_feature, targets: Dict[str, torch.Tensor],
                  alphas_cumprod: torch.Tensor, *, timesteps: Optional[torch.Tensor] = None,
                  noise: Optional[torch.Tensor] = None, dropout: float = DROPOUT_P) -> Dict[str, torch.Tensor]:
    """TrajectoryHead.forward_train (:520-576).  ``timesteps`` / ``noise`` may be injected (the
    reference draws ``randint(0, 50)`` and ``randn``, :533-534)."""
    bs, dev = ego_query.shape[0], ego_query.device
    anchors = head.plan_anchor.unsqueeze(0).expand(bs, -1, -1, -1)
    if timesteps is None:
        timesteps = torch.randint(0, 50, (bs,), device=dev)
    if noise is None:
        noise = torch.randn(anchors.shape, device=dev)
    ac = alphas_cumprod.to(dev)[timesteps]
    noisy = ac.sqrt()[:, None, None, None] * norm_odo(anchors) + (1 - ac).sqrt()[:, None, None, None] * noise
    pts = denorm_odo(noisy.float().clamp(-1, 1))
    feat = head.plan_anchor_encoder(sine_embed(pts).flatten(-2))
    t_emb = time_embed(head, timesteps).view(bs, 1, -1)
    regs, clss = decoder(head, feat, pts, bev_feature, agents_query, ego_query, t_emb, dropout)
    losses = {f"trajectory_loss_{i}": trajectory_loss(r, c, targets["trajectory"], anchors)
              for i, (r, c) in enumerate(zip(regs, clss))}
    mode = clss[-1].argmax(-1)
    best = torch.gather(regs[-1], 1, mode[:, None, None, None].expand(-1, 1, regs[-1].shape[2], 3)).squeeze(1)
    return {"trajectory": best, "trajectory_loss": sum(losses.values()), "trajectory_loss_dict": losses}
