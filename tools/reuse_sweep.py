import os, sys
import torch
sys.path.insert(0, "/root/repo")
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd); head = head.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(1)
for B in (32, 64, 148, 256, 512, 1024, 2048):
    ego = torch.randn(B, 1, 256, device="cuda", generator=g); agents = torch.randn(B, 30, 256, device="cuda", generator=g)
    bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g); nz = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
    res = {}
    for reuse in (1, 0):
        head.set_option("conv_reuse", reuse)
        for _ in range(5): head(ego, agents, bev, noise=nz)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): head(ego, agents, bev, noise=nz)
        b.record(); torch.cuda.synchronize()
        res[reuse] = a.elapsed_time(b) / 20
    print(f"B={B}: reuse {res[1]*1e3:.0f} us ({B/res[1]*1e3:.0f} scenes/s), off {res[0]*1e3:.0f} us ({B/res[0]*1e3:.0f} scenes/s)", flush=True)
