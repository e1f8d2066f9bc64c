"""Small driver for ncu / stage profiling: python tools/prof_run.py --batch 512 --iters 2 [--stages]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--stages", action="store_true")
ap.add_argument("--nhwc", action="store_true")
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
for _ in range(a.iters):
    out = head(ego, agents, bev, noise=noise, **kw)
torch.cuda.synchronize()
if a.stages:
    head.set_profiling(True)
    for _ in range(3):
        head(ego, agents, bev, noise=noise, **kw)
    prof = head.stage_profile()
    tot = sum(v["ms"] for v in prof.values())
    print(f"B={B} {a.precision} stage profile (ms), total {tot:.4f}")
    for k, v in prof.items():
        print(f"  {k:14s} {v['ms']:9.4f} ms  {v['spans']:3d} spans  {100 * v['ms'] / tot:5.1f}%")
print("ok", float(out["trajectory"].abs().sum()))
