"""CPU: pin oracle/head_oracle.py against the committed outputs of the LIVE reference head
(tests/golden/*.npz, written by oracle/make_golden.py) and pin the restated DDIM table."""
import json
import os

import numpy as np
import pytest
import torch

from diffusiondrive_b200 import synth
from diffusiondrive_b200.config import roll_timesteps
from oracle import head_oracle
from oracle.ddim import DDIMSchedulerRestated

# fp32 re-association tolerance: the oracle issues the same torch ops as the reference, so
# it is normally bit-identical; 2e-5 m leaves room for a different CPU's oneDNN kernels.
TOL_M = 2e-5


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _check(out, z, sl):
    assert np.abs(out["trajectory_modes"].numpy() - z["trajectory_modes"][sl]).max() <= TOL_M
    assert np.abs(out["trajectory"].numpy() - z["trajectory"][sl]).max() <= TOL_M
    assert np.abs(out["trajectory_scores"].numpy() - z["trajectory_scores"][sl]).max() <= 1e-4
    assert (out["mode_idx"].numpy() == z["mode_idx"][sl]).all()


def test_oracle_matches_reference_b1(golden_dir):
    z, meta = _load(golden_dir, "default_b1")
    sd = synth.make_state_dict()
    ft = synth.make_features(1)
    out = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"],
                                   synth.make_noise(1))
    assert out["trajectory"].shape == (1, 8, 3)
    _check(out, z, slice(0, 1))


def test_oracle_matches_reference_b256_prefix(golden_dir):
    z, meta = _load(golden_dir, "default_b256")
    assert meta["batch"] == 256
    sd = synth.make_state_dict()
    n = 12
    ft = synth.make_features(n)
    out = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"],
                                   synth.make_noise(n))
    _check(out, z, slice(0, n))
    # B=1 fixture is scene 0 of the B=256 fixture
    z1, _ = _load(golden_dir, "default_b1")
    assert np.abs(z1["trajectory"][0] - z["trajectory"][0]).max() <= TOL_M


def test_oracle_matches_reference_stress(golden_dir):
    z, meta = _load(golden_dir, "stress_b2")
    sd = synth.make_state_dict(num_layers=4, num_anchors=64)
    ft = synth.make_features(2, bev_h=128, bev_w=128)
    out = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"],
                                   synth.make_noise(2, num_anchors=64), num_layers=4, step_num=3)
    assert out["trajectory_modes"].shape == (2, 64, 8, 3)
    _check(out, z, slice(0, 2))


def test_oracle_fp64_truth_close(golden_dir):
    z, _ = _load(golden_dir, "default_b256")
    sd = synth.make_state_dict()
    ft = synth.make_features(2)
    out = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"],
                                   synth.make_noise(2), dtype=torch.float64)
    assert np.abs(out["trajectory_modes"].float().numpy() - z["trajectory_modes"][:2]).max() < 1e-4


def test_ddim_table_and_step(golden_dir):
    """Known values of the scaled_linear table and the step algebra."""
    s = DDIMSchedulerRestated()
    ac = s.alphas_cumprod
    z = np.load(os.path.join(golden_dir, "ddim_alphas_cumprod.npz"))
    assert np.array_equal(ac[:64].numpy(), z["alphas_cumprod"])
    assert abs(float(ac[8]) - 0.9990021586) < 1e-7
    assert abs(float(ac[9]) - 0.9988772273) < 1e-7
    assert abs(float(ac[10]) - 0.9987493157) < 1e-7
    assert abs(float(ac[8]) ** 0.5 - 0.99950093) < 1e-7
    assert abs((1 - float(ac[8])) ** 0.5 - 0.03158863) < 1e-7
    s.set_timesteps(1000)
    g = torch.Generator().manual_seed(5)
    x0 = torch.rand(3, 20, 8, 2, generator=g) * 3 - 1.5      # some values outside [-1, 1]
    xt = torch.randn(3, 20, 8, 2, generator=g)
    got = s.step(x0, 10, xt).prev_sample
    a_t, a_p = ac[10], ac[9]
    eps = (xt - a_t.sqrt() * x0) / (1 - a_t).sqrt()          # uses the UNclipped x0
    want = a_p.sqrt() * x0.clamp(-1, 1) + (1 - a_p).sqrt() * eps
    assert torch.allclose(got, want, atol=1e-6)
    # t = 0 -> prev < 0 -> final_alpha_cumprod = 1 -> prev_sample = clipped x0
    got0 = s.step(x0, 0, xt).prev_sample
    assert torch.allclose(got0, x0.clamp(-1, 1), atol=1e-6)
    # add_noise linearity
    n = torch.randn(3, 20, 8, 2, generator=g)
    t8 = torch.full((3,), 8, dtype=torch.long)
    y = s.add_noise(x0, n, t8)
    assert torch.allclose(y, ac[8].sqrt() * x0 + (1 - ac[8]).sqrt() * n, atol=1e-6)


def test_roll_timesteps():
    assert roll_timesteps(2).tolist() == [10, 0]
    assert roll_timesteps(3).tolist() == [13, 7, 0]


def test_synth_is_deterministic_and_prefix_stable():
    a = synth.make_features(3)
    b = synth.make_features(2)
    for k in a:
        assert torch.equal(a[k][:2], b[k])
    assert torch.equal(synth.make_noise(3)[:2], synth.make_noise(2))
    s1, s2 = synth.make_state_dict(), synth.make_state_dict()
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    anc = synth.make_anchors(20)
    assert anc.shape == (20, 8, 2) and anc.dtype == np.float32
    frac_out = float((anc[..., 0] > 32).mean())
    assert 0.05 < frac_out < 0.2


@pytest.mark.reference
def test_oracle_bit_matches_live_reference(tmp_path):
    """Build-container only: run the live reference module beside the oracle."""
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle.make_golden import run_reference_default
    sd = synth.make_state_dict()
    p = str(tmp_path / "a.npy")
    np.save(p, sd["plan_anchor"].numpy())
    head, _ = ref_import.build_reference_head(sd, p)
    ft = synth.make_features(2)
    nz = synth.make_noise(2)
    traj, reg, cls = run_reference_default(head, ft, nz)
    out = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
    assert (out["trajectory_modes"] - reg).abs().max() <= TOL_M
    assert (out["trajectory"] - traj).abs().max() <= TOL_M
    assert (out["trajectory_scores"] - cls).abs().max() <= 1e-5


def test_full_batch_scores_fixture_is_consistent(golden_dir):
    """default_b4096_scores.npz (scores / modes of 4096 scenes from the live reference) extends
    default_b256.npz: its first 256 scenes are the same reference outputs."""
    a = np.load(os.path.join(golden_dir, "default_b256.npz"))
    b = np.load(os.path.join(golden_dir, "default_b4096_scores.npz"))
    assert b["trajectory_scores"].shape == (4096, 20) and b["mode_idx"].shape == (4096,)
    assert np.array_equal(b["trajectory_scores"][:256], a["trajectory_scores"])
    assert np.array_equal(b["mode_idx"][:256], a["mode_idx"])
    assert np.array_equal(b["trajectory"][:256], a["trajectory"])
    assert np.array_equal(b["mode_idx"], b["trajectory_scores"].argmax(1))
