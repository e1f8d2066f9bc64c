"""Full DiffusionDrive agent forward (SURVEY.md §8f rows N1 + N2): the caller of the planning head.

``DiffusionDriveAgent.forward(features)`` follows ``V2TransfuserModel.forward``
(navsim/agents/diffusiondrive/transfuser_model_v2.py:98-162):

    camera (B,3,256,1024), lidar (B,1,256,256), status (B,8)
      -> TransFuser backbone: two ResNet-34 branches fused at four scales by small GPT blocks,
         FPN top-down to a (B,64,64,64) BEV map          (transfuser_backbone.py:16-277)
      -> 8x8 key/value tokens + status token, 3-layer query decoder -> ego / agent queries  (:112-146)
      -> cross_bev_feature: bilinear 8x8 -> 64x64 of the tokens, concat with the BEV map,
         Linear(320 -> 256) + ReLU + LayerNorm per pixel                                    (:121-140)
      -> TrajectoryHead (the B200-native head of this package)                              (:150-157)

The backbone and the query decoder are *not* on the accelerated path (out of scope): they
are plain PyTorch modules written for this package with the reference's parameter names, so that a
reference checkpoint loads unchanged and the stage can be timed on the GPU box, where neither
``/root/reference`` nor ``timm`` exists.  The producer of ``cross_bev_feature`` (row N1) is a CUDA
kernel behind the C ABI (``ddh_bev_producer``): it never materialises the upsampled 256-channel map
nor the NHWC -> NCHW permute of :138-140 and hands the head NHWC bf16 (or fp32) directly.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .config import HeadConfig
from .trajectory_head import TrajectoryHead


# ------------------------------------------------------------------ backbone pieces
class _BasicBlock(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                            nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = self.bn2(self.conv2(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)), inplace=True)


class _ResNet34(nn.Module):
    """ResNet-34 feature extractor, parameter names of torchvision / timm ``resnet34``."""
    WIDTHS = (64, 128, 256, 512)
    DEPTHS = (3, 4, 6, 3)

    def __init__(self, in_chans: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_chans, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for i, (w, d) in enumerate(zip(self.WIDTHS, self.DEPTHS)):
            blocks = [_BasicBlock(cin if j == 0 else w, w, 2 if (j == 0 and i > 0) else 1)
                      for j in range(d)]
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
            cin = w

    def stem(self, x):
        return F.relu(self.bn1(self.conv1(x)), inplace=True)

    def stage(self, i: int, x):
        if i == 0:
            x = F.max_pool2d(x, 3, 2, 1)
        return getattr(self, f"layer{i + 1}")(x)


class _SelfAttention(nn.Module):
    def __init__(self, d: int, heads: int):
        super().__init__()
        self.key, self.query, self.value = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.proj = nn.Linear(d, d)
        self.n_head = heads

    def forward(self, x):
        b, t, c = x.shape
        split = lambda y: y.view(b, t, self.n_head, c // self.n_head).transpose(1, 2)
        y = F.scaled_dot_product_attention(split(self.query(x)), split(self.key(x)), split(self.value(x)))
        return self.proj(y.transpose(1, 2).reshape(b, t, c))


class _Block(nn.Module):
    def __init__(self, d: int, heads: int, exp: int):
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.attn = _SelfAttention(d, heads)
        self.mlp = nn.Sequential(nn.Linear(d, exp * d), nn.ReLU(True), nn.Linear(exp * d, d), nn.Identity())

    def forward(self, x):
        x = x + self.attn(self.ln1(x))
        return x + self.mlp(self.ln2(x))


class _FusionGPT(nn.Module):
    """Token mixer over the 8x32 image anchors and 8x8 LiDAR anchors (transfuser_backbone.py:279-361)."""

    def __init__(self, d: int, n_img: int, n_lidar: int, heads: int = 4, layers: int = 2, exp: int = 4):
        super().__init__()
        self.pos_emb = nn.Parameter(torch.zeros(1, n_img + n_lidar, d))
        self.blocks = nn.Sequential(*[_Block(d, heads, exp) for _ in range(layers)])
        self.ln_f = nn.LayerNorm(d)
        self.n_img = n_img

    def forward(self, img, lidar):
        b, c, ih, iw = img.shape
        lh, lw = lidar.shape[2:]
        tok = torch.cat((img.permute(0, 2, 3, 1).reshape(b, -1, c), lidar.permute(0, 2, 3, 1).reshape(b, -1, c)), 1)
        x = self.ln_f(self.blocks(self.pos_emb + tok))
        return (x[:, :self.n_img].view(b, ih, iw, c).permute(0, 3, 1, 2),
                x[:, self.n_img:].view(b, lh, lw, c).permute(0, 3, 1, 2))


class TransfuserBackboneTorch(nn.Module):
    """Image + LiDAR fusion backbone (transfuser_backbone.py:16-277) for the default
    ``TransfuserConfig`` (resnet34 x 2, transformer_decoder_join, BEV semantic FPN)."""

    def __init__(self, lidar_chans: int = 1, img_anchors=(8, 32), lidar_anchors=(8, 8), bev_channels: int = 64,
                 lidar_res=(256, 256), bev_down: int = 4):
        super().__init__()
        self.image_encoder = _ResNet34(3)
        self.lidar_encoder = _ResNet34(lidar_chans)
        self.img_anchors, self.lidar_anchors = img_anchors, lidar_anchors
        w = _ResNet34.WIDTHS
        self.transformers = nn.ModuleList([
            _FusionGPT(w[i], img_anchors[0] * img_anchors[1], lidar_anchors[0] * lidar_anchors[1]) for i in range(4)])
        self.lidar_channel_to_img = nn.ModuleList([nn.Conv2d(w[i], w[i], 1) for i in range(4)])
        self.img_channel_to_lidar = nn.ModuleList([nn.Conv2d(w[i], w[i], 1) for i in range(4)])
        self.up_conv5 = nn.Conv2d(bev_channels, bev_channels, 3, padding=1)
        self.up_conv4 = nn.Conv2d(bev_channels, bev_channels, 3, padding=1)
        self.c5_conv = nn.Conv2d(w[3], bev_channels, 1)
        self.p3_size = (lidar_res[0] // bev_down, lidar_res[1] // bev_down)

    def forward(self, image, lidar):
        x, y = self.image_encoder.stem(image), self.lidar_encoder.stem(lidar)
        for i in range(4):
            x, y = self.image_encoder.stage(i, x), self.lidar_encoder.stage(i, y)
            xi = F.adaptive_avg_pool2d(x, self.img_anchors)
            yi = self.lidar_channel_to_img[i](F.adaptive_avg_pool2d(y, self.lidar_anchors))
            xo, yo = self.transformers[i](xi, yi)
            yo = self.img_channel_to_lidar[i](yo)
            x = x + F.interpolate(xo, size=x.shape[2:], mode="bilinear", align_corners=False)
            y = y + F.interpolate(yo, size=y.shape[2:], mode="bilinear", align_corners=False)
        p5 = F.relu(self.c5_conv(y))
        p4 = F.relu(self.up_conv5(F.interpolate(p5, scale_factor=2, mode="bilinear", align_corners=False)))
        p3 = F.relu(self.up_conv4(F.interpolate(p4, size=self.p3_size, mode="bilinear", align_corners=False)))
        return p3, y


class _AgentHead(nn.Module):
    """Bounding-box head (transfuser_model_v2.py:165-205)."""

    def __init__(self, d_ffn: int, d_model: int):
        super().__init__()
        self._mlp_states = nn.Sequential(nn.Linear(d_model, d_ffn), nn.ReLU(), nn.Linear(d_ffn, 5))
        self._mlp_label = nn.Sequential(nn.Linear(d_model, 1))

    def forward(self, q):
        s = self._mlp_states(q)
        s = torch.cat((s[..., :2].tanh() * 32, s[..., 2:3].tanh() * math.pi, s[..., 3:]), -1)
        return {"agent_states": s, "agent_labels": self._mlp_label(q).squeeze(-1)}


# ------------------------------------------------------------------ producer (row N1)
def bev_producer_reference(keyval_tokens, bev_map, weight, bias, ln_w, ln_b):
    """The reference ops of transfuser_model_v2.py:121-140 (returns NCHW fp32): test oracle of the
    CUDA producer and the CPU path of the container-only parity test."""
    b, _, d = keyval_tokens.shape
    g = int(round(math.sqrt(keyval_tokens.shape[1])))
    grid = keyval_tokens.permute(0, 2, 1).contiguous().view(b, d, g, g)
    up = F.interpolate(grid, size=bev_map.shape[2:], mode="bilinear", align_corners=False)
    x = torch.cat((up, bev_map), 1).flatten(-2, -1).permute(0, 2, 1)
    x = F.layer_norm(F.relu(F.linear(x, weight, bias)), (weight.shape[0],), ln_w, ln_b)
    return x.permute(0, 2, 1).contiguous().view(b, -1, *bev_map.shape[2:])


def bev_producer_cuda(keyval_tokens, bev_map, weight, bias, ln_w, ln_b, out_dtype=torch.bfloat16):
    """cross_bev_feature as NHWC (B,H,W,256) through the C ABI (ddh_bev_producer)."""
    lib = _lib.load()
    b, n_tok, d = keyval_tokens.shape
    g = int(round(math.sqrt(n_tok)))
    _, cb, h, w = bev_map.shape
    for t in (keyval_tokens, bev_map, weight, bias, ln_w, ln_b):
        if t.device.type != "cuda" or t.dtype != torch.float32:
            raise RuntimeError("bev_producer_cuda: float32 CUDA tensors expected (there is no CPU fallback)")
    out = torch.empty((b, h, w, weight.shape[0]), dtype=out_dtype, device=bev_map.device)
    scratch = torch.empty((lib.ddh_bev_producer_scratch_bytes(b, g, cb) + 3) // 4, dtype=torch.float32,
                          device=bev_map.device)
    tok, mp, wt = keyval_tokens.contiguous(), bev_map.contiguous(), weight.contiguous()
    rc = lib.ddh_bev_producer(tok.data_ptr(), mp.data_ptr(), wt.data_ptr(), bias.contiguous().data_ptr(),
                              ln_w.contiguous().data_ptr(), ln_b.contiguous().data_ptr(), out.data_ptr(),
                              _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32, b, h, w, g, cb,
                              scratch.data_ptr(),
                              C.c_void_p(torch.cuda.current_stream(bev_map.device).cuda_stream))
    if rc:
        msg = lib.ddh_last_error(None)
        raise RuntimeError(f"ddh_bev_producer failed (ddh_status {rc}): {msg.decode() if msg else ''}")
    return out


# ------------------------------------------------------------------ query decoder (row N3)
class QueryDecoderNative:
    """``_tf_decoder`` (3 x nn.TransformerDecoderLayer, transfuser_model_v2.py:73-80,141-146) and
    ``_agent_head`` (:165-205) behind the C ABI (``ddh_qdec_*``): the head's GEMM engines plus a small
    attention kernel.  Holds no parameters of its own: it packs the agent's modules."""

    def __init__(self, agent: "DiffusionDriveAgent", precision: str):
        self.agent, self.precision = agent, precision
        self._lib = _lib.load()
        self._h = None
        self._sig = None
        self.last_launches = 0

    def __del__(self):
        try:
            if self._h is not None:
                self._lib.ddh_qdec_destroy(self._h)
        except Exception:
            pass

    def _params(self):
        a = self.agent
        return list(a._tf_decoder.parameters()) + list(a._agent_head.parameters()) + [a._query_embedding.weight]

    def _ensure(self, n_keys: int, dev):
        a, lib = self.agent, self._lib
        sig = (self.precision, n_keys) + tuple((p.data_ptr(), p._version) for p in self._params())
        if sig == self._sig:
            return
        layers = list(a._tf_decoder.layers)
        if self._h is None:
            l0 = layers[0]
            shp = _lib.QdecShape(a._query_embedding.weight.shape[0], n_keys, l0.linear1.in_features,
                                 l0.linear1.out_features, l0.self_attn.num_heads, len(layers))
            hp = C.c_void_p()
            rc = lib.ddh_qdec_create(C.byref(shp), C.byref(hp))
            if rc:
                raise RuntimeError(f"ddh_qdec_create failed ({rc}): {lib.ddh_qdec_last_error(None).decode()}")
            self._h = hp
        keep = []

        def ptr(t):
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()
        lw = (_lib.QdecLayerWeights * len(layers))()
        for i, l in enumerate(layers):
            for name, t in (("self_in_w", l.self_attn.in_proj_weight), ("self_in_b", l.self_attn.in_proj_bias),
                            ("self_out_w", l.self_attn.out_proj.weight), ("self_out_b", l.self_attn.out_proj.bias),
                            ("cross_in_w", l.multihead_attn.in_proj_weight), ("cross_in_b", l.multihead_attn.in_proj_bias),
                            ("cross_out_w", l.multihead_attn.out_proj.weight), ("cross_out_b", l.multihead_attn.out_proj.bias),
                            ("lin1_w", l.linear1.weight), ("lin1_b", l.linear1.bias), ("lin2_w", l.linear2.weight),
                            ("lin2_b", l.linear2.bias), ("norm1_w", l.norm1.weight), ("norm1_b", l.norm1.bias),
                            ("norm2_w", l.norm2.weight), ("norm2_b", l.norm2.bias), ("norm3_w", l.norm3.weight),
                            ("norm3_b", l.norm3.bias)):
                setattr(lw[i], name, ptr(t))
        wp = _lib.QdecWeightPtrs()
        wp.query_embedding = ptr(a._query_embedding.weight)
        wp.layers = C.cast(lw, C.POINTER(_lib.QdecLayerWeights))
        ah = a._agent_head
        wp.states0_w, wp.states0_b = ptr(ah._mlp_states[0].weight), ptr(ah._mlp_states[0].bias)
        wp.states2_w, wp.states2_b = ptr(ah._mlp_states[2].weight), ptr(ah._mlp_states[2].bias)
        wp.label_w, wp.label_b = ptr(ah._mlp_label[0].weight), ptr(ah._mlp_label[0].bias)
        stream = torch.cuda.current_stream(dev)
        rc = lib.ddh_qdec_pack_weights(self._h, C.byref(wp), _lib.PREC_BF16 if self.precision == "bf16" else _lib.PREC_FP32,
                                       C.c_void_p(stream.cuda_stream))
        if rc:
            raise RuntimeError(f"ddh_qdec_pack_weights failed ({rc}): {lib.ddh_qdec_last_error(self._h).decode()}")
        stream.synchronize()
        self._sig = sig

    def __call__(self, keyval: torch.Tensor):
        """keyval (B,Nk,256) f32 CUDA -> (query_out (B,Q,256), agent_states (B,Q-1,5), agent_labels (B,Q-1))."""
        if keyval.device.type != "cuda":
            raise RuntimeError("QueryDecoderNative: CUDA tensors expected (there is no CPU fallback)")
        kv = keyval.to(torch.float32).contiguous()
        b, nk, d = kv.shape
        self._ensure(nk, kv.device)
        nq = self.agent._query_embedding.weight.shape[0]
        qo = torch.empty((b, nq, d), dtype=torch.float32, device=kv.device)
        st = torch.empty((b, nq - 1, 5), dtype=torch.float32, device=kv.device)
        lb = torch.empty((b, nq - 1), dtype=torch.float32, device=kv.device)
        rc = self._lib.ddh_qdec_forward(self._h, kv.data_ptr(), qo.data_ptr(), st.data_ptr(), lb.data_ptr(), b,
                                        C.c_void_p(torch.cuda.current_stream(kv.device).cuda_stream))
        if rc:
            raise RuntimeError(f"ddh_qdec_forward failed ({rc}): {self._lib.ddh_qdec_last_error(self._h).decode()}")
        return qo, st, lb


# ------------------------------------------------------------------ the agent
class DiffusionDriveAgent(nn.Module):
    """``V2TransfuserModel`` with the B200-native planning head and BEV producer."""

    def __init__(self, plan_anchor: np.ndarray, config=None, precision: str = "bf16",
                 d_model: int = 256, d_ffn: int = 1024, num_boxes: int = 30, tf_layers: int = 3,
                 num_bev_classes: int = 7, bev_channels: int = 64):
        super().__init__()
        config = config if config is not None else HeadConfig()
        self._query_splits = [1, num_boxes]
        self._backbone = TransfuserBackboneTorch(bev_channels=bev_channels)
        self._keyval_embedding = nn.Embedding(8 ** 2 + 1, d_model)
        self._query_embedding = nn.Embedding(sum(self._query_splits), d_model)
        self._bev_downscale = nn.Conv2d(512, d_model, 1)
        self._status_encoding = nn.Linear(4 + 2 + 2, d_model)
        self._bev_semantic_head = nn.Sequential(
            nn.Conv2d(bev_channels, bev_channels, 3, 1, 1), nn.ReLU(inplace=True),
            nn.Conv2d(bev_channels, num_bev_classes, 1), nn.Upsample(size=(128, 256), mode="bilinear",
                                                                     align_corners=False))
        layer = nn.TransformerDecoderLayer(d_model=d_model, nhead=8, dim_feedforward=d_ffn, dropout=0.0,
                                           batch_first=True)
        self._tf_decoder = nn.TransformerDecoder(layer, tf_layers)
        self._agent_head = _AgentHead(d_ffn, d_model)
        self._trajectory_head = TrajectoryHead(8, d_ffn, d_model, None, config, plan_anchor=plan_anchor,
                                               precision=precision)
        self.bev_proj = nn.Sequential(nn.Linear(d_model + bev_channels, d_model), nn.ReLU(inplace=True),
                                      nn.LayerNorm(d_model))
        self.semantic_map = True          # the reference always computes it (aux output)
        self.native_query_decoder = False  # True: _tf_decoder + _agent_head through ddh_qdec_* (row N3)
        self._qdec = []                    # [QueryDecoderNative] (a list keeps it out of the module tree)
        self.backbone_autocast = None     # e.g. torch.bfloat16: run backbone / query decoder under autocast

    def tokens(self, features: Dict[str, torch.Tensor]):
        """Backbone, 8x8 BEV tokens + status token with the key/value embedding (:110-119)."""
        cam, lidar, status = features["camera_feature"], features["lidar_feature"], features["status_feature"]
        bev_up, x4 = self._backbone(cam, lidar)
        tokens = self._bev_downscale(x4).flatten(-2, -1).permute(0, 2, 1)
        status_enc = self._status_encoding(status)
        keyval = torch.cat((tokens, status_enc[:, None]), 1) + self._keyval_embedding.weight[None]
        return bev_up, keyval, status_enc

    def pre_head(self, features: Dict[str, torch.Tensor]):
        """Everything in front of the planning head, up to (but not including) bev_proj."""
        bev_up, keyval, status_enc = self.tokens(features)
        query = self._query_embedding.weight[None].expand(keyval.shape[0], -1, -1)
        ego_q, agents_q = self._tf_decoder(query, keyval).split(self._query_splits, 1)
        return bev_up, keyval, status_enc, ego_q, agents_q

    def forward(self, features: Dict[str, torch.Tensor], targets=None, *,
                noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        cast = self.backbone_autocast
        native = self.native_query_decoder
        with torch.autocast("cuda", dtype=cast if cast is not None else torch.bfloat16, enabled=cast is not None):
            if native:
                bev_up, keyval, status_enc = self.tokens(features)
            else:
                bev_up, keyval, status_enc, ego_q, agents_q = self.pre_head(features)
            out = {}
            if self.semantic_map:
                out["bev_semantic_map"] = self._bev_semantic_head(bev_up)
        agent_out = None
        if native:
            prec = self._trajectory_head.precision
            if not self._qdec or self._qdec[0].precision != prec:
                self._qdec[:] = [QueryDecoderNative(self, prec)]
            q_out, states, labels = self._qdec[0](keyval.float())
            ego_q, agents_q = q_out.split(self._query_splits, 1)
            agent_out = {"agent_states": states, "agent_labels": labels}
        fp32 = self._trajectory_head.precision == "fp32"
        cross = bev_producer_cuda(keyval[:, :-1].float(), bev_up.float(), self.bev_proj[0].weight,
                                  self.bev_proj[0].bias, self.bev_proj[2].weight, self.bev_proj[2].bias,
                                  torch.float32 if fp32 else torch.bfloat16)
        out.update(self._trajectory_head(ego_q.float().contiguous(), agents_q.float().contiguous(), cross,
                                         tuple(bev_up.shape[2:]), status_enc[:, None], noise=noise,
                                         bev_layout="NHWC"))
        out.update(agent_out if agent_out is not None else self._agent_head(agents_q.float()))
        return out
