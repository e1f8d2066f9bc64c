"""Stage timeline of the cooperative layer kernel at batch 1 (needs tools/_tl/_ddh_tl.so)."""
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
ft = synth.make_features(1); nz = synth.make_noise(1).cuda()
args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
for _ in range(4):
    head(*args, noise=nz)
torch.cuda.synchronize()
d = head.debug_tap("dbg", np.int64)
names = ["bev_out", "sync", "qattn", "sync", "attn_out", "sync", "ffn0", "sync", "ffn2", "sync", "reg0/cls0", "sync", "reg2/cls3", "sync", "tail/cls"]
for cta, off in ((0, 0), (100, 40)):
    t = d[off:off + 16]
    print(f"CTA {cta}: total {int(t[15] - t[0])} cycles")
    for i, n in enumerate(names):
        print(f"   {n:10s} {int(t[i + 1] - t[i]):8d}")
import ctypes as C
lib = _lib.load()
buf = (C.c_longlong * 64)()
lib.ddh_lat_dbg_read.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.ddh_lat_dbg_read(buf))
t = list(buf)[:8]
lab = {(0, 1): "weights+vectors issued", (1, 2): "rows staged, weights to smem", (2, 3): "sync1", (3, 4): "prologue (8 thr/row)", (4, 5): "sync2", (5, 7): "thread-per-output dot"}
print("ffn0 item in CTA 0 (cycles):")
for (a, b), n in lab.items():
    print(f"   {n:30s} {t[b]-t[a]:8d}")
