"""e2e (host buffers) throughput vs host_zero_copy option; B=1 host latency."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth  # noqa: E402
sd = synth.make_state_dict()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator().manual_seed(1)
ego = torch.randn(B, 1, 256, generator=g).pin_memory(); agents = torch.randn(B, 30, 256, generator=g).pin_memory()
bev = torch.randn(B, 256, 64, 64).pin_memory(); noise = torch.randn(B, 20, 8, 2, generator=g).pin_memory()
for zc, hs in ((1, 32), (1, 64)):
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
    head.load_state_dict(sd); head = head.cuda().eval()
    head.set_option("host_zero_copy", zc)
    head.set_option("host_segment", hs)
    for _ in range(2):
        o = head(ego, agents, bev, noise=noise)
    t0 = time.perf_counter()
    for _ in range(5):
        o = head(ego, agents, bev, noise=noise)
        _ = float(o["trajectory"][0, 0, 0])
    dt = (time.perf_counter() - t0) / 5
    w = []
    for _ in range(60):
        t1 = time.perf_counter()
        o1 = head(ego[:1], agents[:1], bev[:1], noise=noise[:1]); _ = o1["trajectory"].numpy()
        w.append((time.perf_counter() - t1) * 1e6)
    w.sort()
    print(json.dumps({"host_zero_copy": zc, "host_segment": hs, "B": B, "ms": dt * 1e3, "scenes_per_s": B / dt, "b1_host_p50_us": w[30]}), flush=True)
