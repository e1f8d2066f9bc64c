"""Conv-kernel timeline (needs a -DDDH_TIMELINE build at tools/_tl/_ddh_tl.so). GPU box only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "_tl", "_ddh_tl.so")
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd); head = head.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
g = torch.Generator(device="cuda").manual_seed(3000)
ego = torch.randn(B, 1, 256, device="cuda", generator=g); agents = torch.randn(B, 30, 256, device="cuda", generator=g)
bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g); noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
for _ in range(3):
    head(ego, agents, bev, noise=noise)
torch.cuda.synchronize()
d = head.debug_tap("dbg", np.int64).reshape(4, 40, 2)
nu = head.debug_tap("nuniq", np.int32)[0]
t0 = d[0, 0, 0]
print("nuniq[0] =", nu)
print("chunk | prod: empty_ok issued | tma: empty_ok issued | mma: full_ok committed   (cycles since first producer stamp)")
for k in range(38):
    r = [int(d[role, k, w] - t0) if d[role, k, w] else -1 for role in range(3) for w in range(2)]
    print(f"{k:3d}   | {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d}")
for i in range(4):
    print("epi", i, [int(x - t0) if x else -1 for x in d[3, i]])
