"""Batch-1 latency probe: eager stream launches vs one CUDA graph replay (GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision=prec)
head.load_state_dict(sd); head = head.cuda().eval()
ft = synth.make_features(B); nz = synth.make_noise(B).cuda()
args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
for _ in range(5):
    out = head(*args, noise=nz)
torch.cuda.synchronize()
head.frozen = True

def timeit(fn, n=200):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); fn(); b.record(); b.synchronize()
    d = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    return d[len(d) // 2], d[int(len(d) * 0.9)]

print(f"{prec} B={B} eager: p50 %.1f us p90 %.1f us, launches {head.last_launch_count()}" % timeit(lambda: head(*args, noise=nz)))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        head(*args, noise=nz)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        gout = head(*args, noise=nz)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
same = all(torch.equal(gout[k], out[k]) for k in out)
print(f"{prec} B={B} graph: p50 %.1f us p90 %.1f us" % timeit(g.replay), "identical:", same)
