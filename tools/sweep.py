"""Timing sweep over scene-chunk concurrency settings (GPU box).
python tools/sweep.py --batch 4096 --chunks 1 2 4 8 [--nhwc]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--chunks", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--nhwc", action="store_true")
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(),
                      precision=a.precision)
head.load_state_dict(sd)
head = head.cuda().eval()
B = a.batch
g = torch.Generator(device="cuda").manual_seed(3000)
ego = torch.randn(B, 1, 256, device="cuda", generator=g)
agents = torch.randn(B, 30, 256, device="cuda", generator=g)
bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
kw = {}
if a.nhwc:
    bev = bev.permute(0, 2, 3, 1).contiguous().bfloat16()
    kw["bev_layout"] = "NHWC"
head(ego, agents, bev, noise=noise, **kw)
ref = None
for c in a.chunks:
    head.set_concurrency(c, 1)
    for _ in range(3):
        out = head(ego, agents, bev, noise=noise, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = head(ego, agents, bev, noise=noise, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    same = True if ref is None else all(torch.equal(out[k], ref[k]) for k in out)
    if ref is None:
        ref = {k: v.clone() for k, v in out.items()}
    print(f"chunks={c:2d}  {ms:8.3f} ms/step  {B / ms * 1e3:10.0f} scenes/s  launches={head.last_launch_count()}"
          f"  identical_to_first={same}", flush=True)
