"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the LIVE reference head.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For every fixture the synthetic weights / anchors / features / noise are
regenerated from seeds (diffusiondrive_b200/synth.py), loaded into the reference
``TrajectoryHead`` (imported through oracle/ref_import.py), and the reference's
own outputs are stored: ``trajectory`` (the eval return value,
transfuser_model_v2.py:641) and the locals ``poses_reg`` / ``poses_cls`` of the
last decoder call (:630-631), captured with a forward hook on ``diff_decoder``.
Inputs are NOT stored: the tests regenerate them from the same seeds.

The stress fixture (64 anchors, 3 denoise steps, 4 decoder layers, 128x128 BEV)
drives the reference's own sub-modules from a harness loop because ``step_num``
is a local literal of ``forward_test`` (:581).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from diffusiondrive_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class _InjectNoise:
    """Replace ``torch.randn`` for the duration of one reference forward (:593)."""

    def __init__(self, noise):
        self.noise = noise

    def __enter__(self):
        self._orig = torch.randn
        noise = self.noise

        def fake(*shape, **kw):
            shp = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
            assert tuple(shp) == tuple(noise.shape), (shp, noise.shape)
            return noise.clone()
        torch.randn = fake
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        return False


def run_reference_default(head, feats, noise):
    cap = {}
    h = head.diff_decoder.register_forward_hook(
        lambda m, i, o: cap.update(reg=o[0][-1].clone(), cls=o[1][-1].clone()))
    try:
        with torch.no_grad(), _InjectNoise(noise):
            out = head(feats["ego_query"], feats["agents_query"], feats["bev_feature"],
                       tuple(feats["bev_feature"].shape[2:]), feats["status_encoding"])
    finally:
        h.remove()
    return out["trajectory"], cap["reg"], cap["cls"]


def run_reference_steps(head, mod, feats, noise, step_num):
    """forward_test body (:578-641) re-driven with the reference's own sub-modules for a
    non-default ``step_num``."""
    from navsim.agents.diffusiondrive.modules.blocks import gen_sineembed_for_position
    ego, agents, bev = feats["ego_query"], feats["agents_query"], feats["bev_feature"]
    bs = ego.shape[0]
    sch = head.diffusion_scheduler
    with torch.no_grad():
        sch.set_timesteps(1000, ego.device)
        roll = (np.arange(0, step_num) * (20 / step_num)).round()[::-1].copy().astype(np.int64)
        roll = torch.from_numpy(roll)
        plan_anchor = head.plan_anchor.unsqueeze(0).repeat(bs, 1, 1, 1)
        img = head.norm_odo(plan_anchor)
        trunc = torch.ones((bs,), dtype=torch.long) * 8
        img = sch.add_noise(original_samples=img, noise=noise, timesteps=trunc)
        modes = img.shape[1]
        for k in roll[:]:
            x_boxes = torch.clamp(img, min=-1, max=1)
            pts = head.denorm_odo(x_boxes)
            emb = gen_sineembed_for_position(pts, hidden_dim=64).flatten(-2)
            f = head.plan_anchor_encoder(emb).view(bs, modes, -1)
            ts = k[None].expand(bs)
            te = head.time_mlp(ts).view(bs, 1, -1)
            regs, clss = head.diff_decoder(f, pts, bev, tuple(bev.shape[2:]), agents, ego, te,
                                           feats["status_encoding"], None)
            reg, cls = regs[-1], clss[-1]
            x_start = head.norm_odo(reg[..., :2])
            img = sch.step(model_output=x_start, timestep=k, sample=img).prev_sample
        mode_idx = cls.argmax(dim=-1)
        gi = mode_idx[..., None, None, None].repeat(1, 1, 8, 3)
        best = torch.gather(reg, 1, gi).squeeze(1)
    return best, reg, cls


def _save(name, traj, reg, cls, meta):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, trajectory=traj.numpy().astype(np.float32),
                        trajectory_modes=reg.numpy().astype(np.float32),
                        trajectory_scores=cls.numpy().astype(np.float32),
                        mode_idx=cls.argmax(-1).numpy().astype(np.int64),
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


def make_agent_fixture(tmp):
    """BASELINE configs[3]: the LIVE reference V2TransfuserModel (timm stand-in on torchvision's
    resnet34, oracle/ref_import.py) on seeded synthetic weights / sensors, batch 1: the planned
    trajectory, all-mode poses and scores, the queries and a subsample of cross_bev_feature."""
    from diffusiondrive_b200.agent import DiffusionDriveAgent
    anchors = synth.make_state_dict()["plan_anchor"].numpy()
    mine = DiffusionDriveAgent(anchors, precision="fp32").eval()      # (only to enumerate tensor names)
    sd = synth.make_agent_state_dict(mine)
    apath = os.path.join(tmp, "anchors20_agent.npy")
    np.save(apath, anchors)
    ref, _cfg = ref_import.build_reference_model(sd, apath)
    feats = synth.make_agent_inputs(1)
    noise = synth.make_noise(1)
    cap = {}
    h1 = ref._trajectory_head.register_forward_pre_hook(
        lambda m, a: cap.update(ego=a[0].clone(), agents=a[1].clone(), bev=a[2].clone()))
    h2 = ref._trajectory_head.diff_decoder.register_forward_hook(
        lambda m, i, o: cap.update(reg=o[0][-1].clone(), cls=o[1][-1].clone()))
    try:
        with torch.no_grad(), _InjectNoise(noise):
            out = ref(feats)
    finally:
        h1.remove(), h2.remove()
    path = os.path.join(GOLDEN_DIR, "full_agent_b1.npz")
    np.savez_compressed(
        path, trajectory=out["trajectory"].numpy(), trajectory_modes=cap["reg"].numpy(),
        trajectory_scores=cap["cls"].numpy(), mode_idx=cap["cls"].argmax(-1).numpy().astype(np.int64),
        agent_states=out["agent_states"].numpy(), agent_labels=out["agent_labels"].numpy(),
        ego_query=cap["ego"].numpy(), agents_query=cap["agents"].numpy(),
        cross_bev_sub=cap["bev"][:, :, ::8, ::8].numpy(), cross_bev_abs_mean=np.float32(cap["bev"].abs().mean()),
        bev_semantic_sub=out["bev_semantic_map"][:, :, ::16, ::16].numpy(),
        meta=np.frombuffer(json.dumps({"torch": torch.__version__, "seed_agent": synth.SEED_AGENT,
                                       "gain": synth.AGENT_GAIN, "params": int(sum(p.numel() for p in ref.parameters())),
                                       "source": "live reference V2TransfuserModel via oracle/ref_import.py"}).encode(),
                           dtype=np.uint8))
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    tmp = tempfile.mkdtemp()
    if "--agent-only" in sys.argv:
        make_agent_fixture(tmp)
        return
    meta_base = {"torch": torch.__version__, "seed_weights": synth.SEED_WEIGHTS,
                 "seed_features": synth.SEED_FEATURES, "seed_noise": synth.SEED_NOISE,
                 "source": "live reference TrajectoryHead via oracle/ref_import.py"}

    # ---- default configuration: 20 anchors, 2 steps, 2 layers, 64x64 BEV
    sd = synth.make_state_dict()
    apath = os.path.join(tmp, "anchors20.npy")
    np.save(apath, sd["plan_anchor"].numpy())
    head, _cfg = ref_import.build_reference_head(sd, apath, num_layers=2)
    for name, B in (("default_b1", 1), ("default_b256", 256)):
        trajs, regs, clss = [], [], []
        for s0 in range(0, B, 32):
            n = min(32, B - s0)
            feats = synth.make_features(n, start=s0)
            noise = synth.make_noise(n, start=s0)
            t, r, c = run_reference_default(head, feats, noise)
            trajs.append(t), regs.append(r), clss.append(c)
        _save(name, torch.cat(trajs), torch.cat(regs), torch.cat(clss),
              dict(meta_base, batch=B, anchors=20, steps=2, layers=2, bev=[64, 64]))

    # ---- mode-agreement statistics at BASELINE's full batch: scores / selected mode / selected
    # trajectory of 4096 scenes (all-mode poses are not stored: 7.9 MB)
    if "--full" in sys.argv:
        B = 4096
        trajs, clss = [], []
        for s0 in range(0, B, 32):
            feats = synth.make_features(32, start=s0)
            noise = synth.make_noise(32, start=s0)
            t, _r, c = run_reference_default(head, feats, noise)
            trajs.append(t), clss.append(c)
        cls = torch.cat(clss)
        path = os.path.join(GOLDEN_DIR, "default_b4096_scores.npz")
        np.savez_compressed(path, trajectory=torch.cat(trajs).numpy().astype(np.float32),
                            trajectory_scores=cls.numpy().astype(np.float32),
                            mode_idx=cls.argmax(-1).numpy().astype(np.int64),
                            meta=np.frombuffer(json.dumps(dict(meta_base, batch=B, anchors=20, steps=2,
                                                               layers=2, bev=[64, 64])).encode(), dtype=np.uint8))
        print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")

    # ---- stress configuration: 64 anchors, 3 steps, 4 layers, 128x128 BEV
    sd = synth.make_state_dict(num_layers=4, num_anchors=64)
    apath = os.path.join(tmp, "anchors64.npy")
    np.save(apath, sd["plan_anchor"].numpy())
    head, _cfg = ref_import.build_reference_head(sd, apath, num_layers=4)
    _, _, mod = ref_import.load_reference()
    B = 2
    feats = synth.make_features(B, bev_h=128, bev_w=128)
    noise = synth.make_noise(B, num_anchors=64)
    t, r, c = run_reference_steps(head, mod, feats, noise, step_num=3)
    _save("stress_b2", t, r, c,
          dict(meta_base, batch=B, anchors=64, steps=3, layers=4, bev=[128, 128]))

    make_agent_fixture(tmp)

    # ---- DDIM table pin (oracle/ddim.py is a restatement: "parity unpinned" upstream)
    from oracle.ddim import DDIMSchedulerRestated
    ac = DDIMSchedulerRestated().alphas_cumprod[:64].numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ddim_alphas_cumprod.npz"), alphas_cumprod=ac)


if __name__ == "__main__":
    main()
