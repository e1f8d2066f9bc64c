"""Dense-GEMM timeline (needs tools/_tl/_ddh_tl.so, env DDH_TIMELINE_GEMM=<launch idx>). GPU box only."""
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(3000)
ego = torch.randn(B, 1, 256, device="cuda", generator=g); agents = torch.randn(B, 30, 256, device="cuda", generator=g)
bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g); noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
head(ego, agents, bev, noise=noise)
head.set_concurrency(1, 1)
for _ in range(2):
    head(ego, agents, bev, noise=noise)
torch.cuda.synchronize()
d = head.debug_tap("dbg", np.int64).reshape(4, 40, 2)
t0 = d[2, 0, 0]
print("launch idx", os.environ.get("DDH_TIMELINE_GEMM"), "B", B)
print("tile | epi: wait_start | epi: tfull_ok  done | mma: tempty_ok committed | mma: first_full_ok")
for k in range(7):
    f = lambda x: int(x - t0) if x else -1
    print(f"{k:3d}  | {f(d[0,k,0]):8d} | {f(d[1,k,0]):8d} {f(d[1,k,1]):8d} | {f(d[2,k,0]):8d} {f(d[2,k,1]):8d} | {f(d[3,k,0]):8d}")
print("pass-1 blocks of tile 1 (cycles rel. to MMA start): ld_start ld_done | commit_start commit_done | fin_start fin_done")
for b in range(8):
    f = lambda x: int(x - t0) if x else -1
    print(b, f(d[3, 8 + b*4, 0]), f(d[3, 8 + b*4, 1]), "|", f(d[3, 9 + b*4, 0]), f(d[3, 9 + b*4, 1]), "|", f(d[3, 10 + b*4, 0]), f(d[3, 10 + b*4, 1]))
