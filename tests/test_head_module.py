"""CPU: the drop-in module mirrors the reference's parameter tree and interface, and never
falls back to a CPU computation."""
import numpy as np
import pytest
import torch

from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth


def _head(**kw):
    sd = synth.make_state_dict()
    return TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(),
                          **kw), sd


def test_state_dict_names_match_reference():
    head, sd = _head()
    own = head.state_dict()
    assert set(own) == set(sd)
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    assert sum(p.numel() for p in head.parameters()) == 4_950_658
    res = head.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert not head.plan_anchor.requires_grad


def test_reference_default_inits():
    head, _ = _head()
    l0, l1 = head.diff_decoder.layers[0], head.diff_decoder.layers[1]
    assert float(l0.cross_bev_attention.attention_weights.weight.abs().max()) == 0.0
    assert abs(float(l0.task_decoder.plan_cls_branch[-1].bias) + 4.59512) < 1e-4
    # _get_clones semantics: layers start identical
    assert torch.equal(l0.ffn[0].weight, l1.ffn[0].weight)


def test_anchor_file_constructor(tmp_path):
    p = tmp_path / "anchors.npy"
    np.save(p, synth.make_anchors(20))
    head = TrajectoryHead(num_poses=8, d_ffn=1024, d_model=256, plan_anchor_path=str(p),
                          config=HeadConfig())
    assert tuple(head.plan_anchor.shape) == (20, 8, 2)


def test_no_cpu_fallback_for_inference():
    head, sd = _head()
    ft = synth.make_features(1)
    head.train()
    with pytest.raises(ValueError, match="targets"):      # training mode needs targets (forward_train)
        head(ft["ego_query"], ft["agents_query"], ft["bev_feature"], (64, 64), ft["status_encoding"])
    head.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head(ft["ego_query"], ft["agents_query"], ft["bev_feature"], (64, 64), ft["status_encoding"])
    with pytest.raises(ValueError):
        TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=synth.make_anchors(20),
                       precision="fp8")
