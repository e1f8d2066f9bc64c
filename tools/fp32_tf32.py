"""fp32 engine with the 3xTF32 tensor-core conv (option fp32_tensor_conv) against the CUDA-core conv and the
live-reference golden: parity and time.   [CHECKED=1] python tools/fp32_tf32.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402
if os.environ.get("CHECKED"):
    _lib.use_library(os.path.join(ROOT, "diffusiondrive_b200", "_ddh_checked.so"))
B = 256
z = np.load(os.path.join(ROOT, "tests", "golden", "default_b256.npz"))
sd = synth.make_state_dict()
ft = synth.make_features(B)
nz = synth.make_noise(B).cuda()
args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
outs = {}
for tc in (1, 0):
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="fp32")
    head.load_state_dict(sd)
    head = head.cuda().eval()
    head.set_option("fp32_tensor_conv", tc)
    for _ in range(2):
        o = head(*args, noise=nz)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        o = head(*args, noise=nz)
    b.record()
    torch.cuda.synchronize()
    m = o["trajectory_modes"].cpu().numpy()
    dxy = np.abs(m[..., :2] - z["trajectory_modes"][..., :2]).max()
    dh = np.abs(m[..., 2] - z["trajectory_modes"][..., 2]).max()
    ds = np.abs(o["trajectory_scores"].cpu().numpy() - z["trajectory_scores"]).max()
    agree = (o["mode_idx"].cpu().numpy() == z["mode_idx"]).mean()
    outs[tc] = m
    print(f"fp32_tensor_conv={tc}: {a.elapsed_time(b) / 3:.3f} ms per forward of {B} scenes ({B / (a.elapsed_time(b) / 3) * 1e3:.0f} scenes/s), "
          f"launches {head.last_launch_count()}, vs golden: xy {dxy:.3e} m, heading {dh:.3e}, score {ds:.3e}, modes agree {agree:.4f}", flush=True)
print(f"tensor vs CUDA-core conv: max |d| {np.abs(outs[1] - outs[0]).max():.3e}")
