"""A handful of batch-1 forwards on the resident engine (ncu target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd); head = head.cuda().eval()
ft = synth.make_features(B); nz = synth.make_noise(B).cuda()
args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
for _ in range(n):
    out = head(*args, noise=nz)
torch.cuda.synchronize()
print("launches", head.last_launch_count(), "mode", out["mode_idx"].tolist())
