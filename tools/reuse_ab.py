"""A/B of the value-row reuse across denoise steps (option conv_reuse): outputs, conv rows, stage times.
[CHECKED=1] python tools/reuse_ab.py [B]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402
if os.environ.get("CHECKED"):
    _lib.use_library(os.path.join(ROOT, "diffusiondrive_b200", "_ddh_checked.so"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd)
head = head.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(3000)
n_par = min(B, 256)
ft = synth.make_features(n_par)
ego = torch.randn(B, 1, 256, device="cuda", generator=g)
agents = torch.randn(B, 30, 256, device="cuda", generator=g)
bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
ego[:n_par] = ft["ego_query"].cuda(); agents[:n_par] = ft["agents_query"].cuda()
bev[:n_par] = ft["bev_feature"].cuda(); noise[:n_par] = synth.make_noise(n_par).cuda()
outs = {}
for reuse in (1, 0, 1, 0):
    head.set_option("conv_reuse", reuse)
    for _ in range(3):
        out = head(ego, agents, bev, noise=noise)
    torch.cuda.synchronize()
    outs.setdefault(reuse, {k: v.clone() for k, v in out.items()})
    head.set_profiling(True)
    head(ego, agents, bev, noise=noise)
    prof = head.stage_profile()
    rows = head.debug_tap("conv_rows", np.int32)
    head.set_profiling(False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        head(ego, agents, bev, noise=noise)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"conv_reuse {reuse}: step {ms:.3f} ms ({B / ms * 1e3:.0f} scenes/s), launches {head.last_launch_count()}, "
          f"rows per call {rows.tolist()}", flush=True)
    print("   " + ", ".join(f"{k} {v['ms']:.3f}/{v['spans']}" for k, v in prof.items() if v["spans"]), flush=True)
a, b = outs[1], outs[0]
d = (a["trajectory_modes"] - b["trajectory_modes"])[..., :2].abs().max().item()
print(f"reuse vs no reuse: max |dxy| {d:.3e} m, mode agreement {(a['mode_idx'] == b['mode_idx']).float().mean().item():.5f}, "
      f"max |dscore| {(a['trajectory_scores'] - b['trajectory_scores']).abs().max().item():.3e}")
gp = os.path.join(ROOT, "tests", "golden", "default_b256.npz")
if n_par == 256 and os.path.exists(gp):
    z = np.load(gp)
    for name, o in (("reuse", a), ("no reuse", b)):
        m = o["trajectory_modes"][:256].float().cpu().numpy()
        dxy = np.abs(m[..., :2] - z["trajectory_modes"][..., :2]).max()
        agree = (o["mode_idx"][:256].cpu().numpy() == z["mode_idx"]).mean()
        print(f"{name} vs live-reference golden: max |dxy| {dxy:.3e} m, mode agreement {agree:.4f}")
# rerun determinism with reuse on
head.set_option("conv_reuse", 1)
c = head(ego, agents, bev, noise=noise)
print("rerun bit-identical:", all(torch.equal(c[k], a[k]) for k in a))
