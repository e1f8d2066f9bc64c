"""Training path: TrajectoryHead in training mode delegates to the
differentiable PyTorch restatement of forward_train + LossComputer (train_torch.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, train_torch
from diffusiondrive_b200.trajectory_head import ddim_alphas_cumprod


def _inputs(B):
    ft = synth.make_features(B)
    tgt = {"trajectory": torch.randn(B, 8, 3, generator=torch.Generator().manual_seed(3)) * 5}
    return ft, synth.make_noise(B), tgt


def test_training_mode_returns_loss_and_gradients():
    sd = synth.make_state_dict()
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy())
    head.load_state_dict(sd)
    head.train()
    ft, nz, tgt = _inputs(2)
    torch.manual_seed(0)
    out = head(ft["ego_query"], ft["agents_query"], ft["bev_feature"], (64, 64), ft["status_encoding"],
               targets=tgt, noise=nz)
    assert set(out) == {"trajectory", "trajectory_loss", "trajectory_loss_dict"}
    assert out["trajectory"].shape == (2, 8, 3)
    assert set(out["trajectory_loss_dict"]) == {"trajectory_loss_0", "trajectory_loss_1"}
    loss = out["trajectory_loss"]
    assert torch.isfinite(loss) and loss.requires_grad
    loss.backward()
    g = head.diff_decoder.layers[1].task_decoder.plan_reg_branch[4].weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    assert head.plan_anchor.grad is None        # frozen parameter (:455-458)


def test_forward_train_matches_live_reference():
    """Container only: loss, selected trajectory and every parameter gradient equal the live
    reference's forward_train + LossComputer on the same timesteps / noise (dropout off)."""
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    sd = synth.make_state_dict()
    tmp = tempfile.mkdtemp()
    ap = os.path.join(tmp, "a.npy")
    np.save(ap, sd["plan_anchor"].numpy())
    ref, _cfg = ref_import.build_reference_head(sd, ap)
    ref.train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    B = 3
    ft, nz, tgt = _inputs(B)
    ts = torch.tensor([3, 17, 44])
    orig = (torch.randint, torch.randn)
    torch.randint = lambda *a, **k: ts.clone()
    torch.randn = lambda *a, **k: nz.clone()
    try:
        r = ref(ft["ego_query"], ft["agents_query"], ft["bev_feature"], (64, 64), ft["status_encoding"], targets=tgt)
    finally:
        torch.randint, torch.randn = orig
    r["trajectory_loss"].backward()
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy())
    head.load_state_dict(sd)
    head.train()
    o = train_torch.forward_train(head, ft["ego_query"], ft["agents_query"], ft["bev_feature"], tgt,
                                  ddim_alphas_cumprod(), timesteps=ts, noise=nz, dropout=0.0)
    o["trajectory_loss"].backward()
    assert abs(float(o["trajectory_loss"]) - float(r["trajectory_loss"])) < 1e-4
    assert (o["trajectory"] - r["trajectory"]).abs().max() < 1e-5
    ref_grads = {k: v.grad for k, v in ref.named_parameters() if v.grad is not None}
    n = 0
    for k, v in head.named_parameters():
        if v.grad is not None:
            assert k in ref_grads and (v.grad - ref_grads[k]).abs().max() < 1e-4, k
            n += 1
    assert n == len(ref_grads) and n > 80
