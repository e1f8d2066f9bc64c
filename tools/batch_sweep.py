"""Latency of one forward vs batch size for the bf16 engines (GPU box).
Usage: python tools/batch_sweep.py [resident_engine 0|1]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth

sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd); head = head.cuda().eval()
RES = int(sys.argv[1]) if len(sys.argv) > 1 else 1
head.set_option("resident_engine", RES)
for B in (1, 2, 4, 8, 12, 16, 20, 24, 32, 64):
    ft = synth.make_features(B); nz = synth.make_noise(B).cuda()
    args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
    for _ in range(5):
        head(*args, noise=nz)
    torch.cuda.synchronize()
    head.frozen = True
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for a, b in evs:
        a.record(); head(*args, noise=nz); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    print(f"resident_engine={RES} B={B} launches={head.last_launch_count()} p50 {ts[len(ts)//2]:.1f} us "
          f"({B / ts[len(ts)//2] * 1e6:.0f} scenes/s)", flush=True)
