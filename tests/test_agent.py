"""Full-agent forward (SURVEY.md section 8f rows N1 + N2): the PyTorch backbone / query decoder
written for this package against the LIVE reference V2TransfuserModel (container only), the CUDA
cross_bev_feature producer against the reference ops, and the whole agent on the GPU against the
reference's committed outputs (tests/golden/full_agent_b1.npz)."""
import os

import numpy as np
import pytest
import torch

from diffusiondrive_b200 import synth
from diffusiondrive_b200.agent import (DiffusionDriveAgent, bev_producer_cuda, bev_producer_reference)


def _agent(precision, device="cpu"):
    anchors = synth.make_state_dict()["plan_anchor"].numpy()
    agent = DiffusionDriveAgent(anchors, precision=precision).eval()
    agent.load_state_dict(synth.make_agent_state_dict(agent))
    return agent.to(device)


def test_agent_state_dict_is_reference_compatible():
    agent = _agent("fp32")
    names = set(agent.state_dict())
    for k in ("_backbone.image_encoder.layer4.2.conv2.weight", "_backbone.lidar_encoder.conv1.weight",
              "_backbone.transformers.3.blocks.1.attn.query.weight", "_backbone.c5_conv.bias",
              "_keyval_embedding.weight", "_tf_decoder.layers.2.multihead_attn.in_proj_weight",
              "_agent_head._mlp_states.2.weight", "bev_proj.0.weight", "bev_proj.2.bias",
              "_trajectory_head.diff_decoder.layers.1.task_decoder.plan_reg_branch.4.weight"):
        assert k in names, k
    assert agent.state_dict()["_backbone.lidar_encoder.conv1.weight"].shape == (64, 1, 7, 7)
    assert sum(p.numel() for p in agent.parameters()) == 60_715_727      # the reference's 60.7 M


def test_pre_head_matches_live_reference(golden_dir):
    """Container only (needs /root/reference): backbone + tokens + query decoder + producer ops of
    this package reproduce the tensors the reference hands its TrajectoryHead."""
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    z = np.load(os.path.join(golden_dir, "full_agent_b1.npz"))
    agent = _agent("fp32")
    feats = synth.make_agent_inputs(1)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        bev_up, keyval, _status, ego_q, agents_q = agent.pre_head(feats)
        cross = bev_producer_reference(keyval[:, :-1], bev_up, agent.bev_proj[0].weight, agent.bev_proj[0].bias,
                                       agent.bev_proj[2].weight, agent.bev_proj[2].bias)
        heads = agent._agent_head(agents_q)
    assert np.abs(cross[:, :, ::8, ::8].numpy() - z["cross_bev_sub"]).max() < 1e-4
    assert abs(float(cross.abs().mean()) - float(z["cross_bev_abs_mean"])) < 1e-5
    assert np.abs(ego_q.numpy() - z["ego_query"]).max() < 2e-4
    assert np.abs(agents_q.numpy() - z["agents_query"]).max() < 2e-4
    assert np.abs(heads["agent_states"].numpy() - z["agent_states"]).max() < 1e-3
    assert np.abs(heads["agent_labels"].numpy() - z["agent_labels"]).max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_bev_producer_kernel_matches_reference_ops(dtype, tol):
    g = torch.Generator().manual_seed(5)
    B = 3
    tok = torch.randn(B, 64, 256, generator=g)
    bev = torch.randn(B, 64, 64, 64, generator=g)
    w = torch.randn(256, 320, generator=g) / 320 ** 0.5
    b = 0.1 * torch.randn(256, generator=g)
    lg = 1 + 0.1 * torch.randn(256, generator=g)
    lb = 0.1 * torch.randn(256, generator=g)
    ref = bev_producer_reference(tok.double(), bev.double(), w.double(), b.double(), lg.double(), lb.double())
    out = bev_producer_cuda(tok.cuda(), bev.cuda(), w.cuda(), b.cuda(), lg.cuda(), lb.cuda(), dtype)
    torch.cuda.synchronize()
    assert out.shape == (B, 64, 64, 256) and out.dtype == dtype
    got = out.float().cpu().permute(0, 3, 1, 2).double()
    if dtype == torch.bfloat16:      # bf16 output: half an ulp of the stored value (2^-9 relative)
        assert ((got - ref).abs() <= ref.abs() * 2.0 ** -8 + 1e-3).all()
    d = (got - ref).abs().max().item()
    assert d < tol * (3 if dtype == torch.bfloat16 else 1), d
    with pytest.raises(RuntimeError):
        bev_producer_cuda(tok, bev, w, b, lg, lb)          # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 6e-2)])
def test_native_query_decoder_matches_torch_transformer_decoder(precision, tol):
    """Row N3: ddh_qdec_* (3-layer post-norm decoder, 31 queries x 65 keys, + AgentHead) against
    torch.nn.TransformerDecoder / the PyTorch AgentHead on the same weights."""
    from diffusiondrive_b200.agent import QueryDecoderNative
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        agent = _agent(precision, "cuda")
        g = torch.Generator().manual_seed(11)
        for B in (1, 5):
            keyval = (2.0 * torch.randn(B, 65, 256, generator=g)).cuda()
            with torch.no_grad():
                query = agent._query_embedding.weight[None].expand(B, -1, -1)
                ref_q = agent._tf_decoder(query, keyval)
                ref_a = agent._agent_head(ref_q[:, 1:])
            q, st, lb = QueryDecoderNative(agent, precision)(keyval)
            torch.cuda.synchronize()
            assert q.shape == (B, 31, 256) and st.shape == (B, 30, 5) and lb.shape == (B, 30)
            assert (q - ref_q).abs().max().item() < tol, (q - ref_q).abs().max().item()
            assert (st - ref_a["agent_states"]).abs().max().item() < tol * 20
            assert (lb - ref_a["agent_labels"]).abs().max().item() < tol * 5
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.gpu
def test_full_agent_matches_reference_golden(golden_dir):
    """BASELINE configs[3] parity: same random weights, sensors and DDIM noise as the live
    reference model; fp32 head within 1e-4 m of the reference trajectory, bf16 within 2e-2 m."""
    z = np.load(os.path.join(golden_dir, "full_agent_b1.npz"))
    feats = {k: v.cuda() for k, v in synth.make_agent_inputs(1).items()}
    noise = synth.make_noise(1).cuda()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for precision, tol, native in (("fp32", 1e-4, False), ("bf16", 2e-2, False), ("fp32", 1e-4, True),
                                       ("bf16", 2e-2, True)):
            agent = _agent(precision, "cuda")
            agent.native_query_decoder = native
            with torch.no_grad():
                out = agent(feats, noise=noise)
            torch.cuda.synchronize()
            dq = np.abs(out["agent_states"].cpu().numpy() - z["agent_states"]).max()
            dm = np.abs(out["trajectory_modes"].cpu().numpy() - z["trajectory_modes"]).max()
            dt = np.abs(out["trajectory"].cpu().numpy() - z["trajectory"]).max()
            print("AGENT PARITY", precision, "native_qdec", native, "modes", dm, "trajectory", dt, "agent_states", dq)
            assert out["trajectory"].shape == (1, 8, 3)
            assert dm <= tol, (precision, dm)
            if int(out["mode_idx"][0]) == int(z["mode_idx"][0]):
                assert dt <= tol
            assert dq < (5e-3 if precision == "fp32" or not native else 0.5)
            assert np.abs(out["bev_semantic_map"][:, :, ::16, ::16].cpu().numpy() - z["bev_semantic_sub"]).max() < 1e-3
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
