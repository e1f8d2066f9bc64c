"""Small end-to-end run of every engine for compute-sanitizer (GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth
sd = synth.make_state_dict()
for prec, B in (("bf16", 1), ("bf16", 7), ("fp32", 3)):
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision=prec)
    head.load_state_dict(sd); head = head.cuda().eval()
    ft = synth.make_features(B); nz = synth.make_noise(B).cuda()
    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(), noise=nz)
    torch.cuda.synchronize()
    print(prec, B, float(out["trajectory"].abs().sum()), head.last_launch_count())
print("done")
