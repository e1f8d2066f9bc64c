"""A/B of the persistent conv's dynamic scene queue (option conv_dynamic): identical results, stage times.
[CHECKED=1] python tools/conv_dyn.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402
if os.environ.get("CHECKED"):
    _lib.use_library(os.path.join(ROOT, "diffusiondrive_b200", "_ddh_checked.so"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd)
head = head.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(3000)
ego = torch.randn(B, 1, 256, device="cuda", generator=g)
agents = torch.randn(B, 30, 256, device="cuda", generator=g)
bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
outs = {}
for dyn in (1, 0, 1, 0):
    head.set_option("conv_dynamic", dyn)
    for _ in range(3):
        out = head(ego, agents, bev, noise=noise)
    torch.cuda.synchronize()
    outs.setdefault(dyn, out)
    head.set_profiling(True)
    for _ in range(3):
        head(ego, agents, bev, noise=noise)
    prof = head.stage_profile()
    head.set_profiling(False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        head(ego, agents, bev, noise=noise)
    b.record()
    torch.cuda.synchronize()
    print(f"conv_dynamic {dyn}: conv {prof['conv']['ms']:.3f} ms, step {a.elapsed_time(b) / 5:.3f} ms "
          f"({B / (a.elapsed_time(b) / 5) * 1e3:.0f} scenes/s)", flush=True)
same = all(torch.equal(outs[1][k], outs[0][k]) for k in outs[1])
print("dynamic == static:", same)
