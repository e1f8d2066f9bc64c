"""Experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 B) vs the on-demand layout pass."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth  # noqa: E402
rt = ctypes.CDLL("libcudart.so.12")
gran = int(sys.argv[1])
torch.cuda.init()
torch.zeros(1, device="cuda")
if gran:
    print("cudaDeviceSetLimit rc", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran)))
v = ctypes.c_size_t()
rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
print("L2 fetch granularity limit", v.value)
sd = synth.make_state_dict()
head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
head.load_state_dict(sd); head = head.cuda().eval()
B = 4096
g = torch.Generator(device="cuda").manual_seed(3000)
ego = torch.randn(B, 1, 256, device="cuda", generator=g); agents = torch.randn(B, 30, 256, device="cuda", generator=g)
bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g); noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
for seg in (8, 16):
    head.set_option("layout_segment", seg)
    for _ in range(3):
        head(ego, agents, bev, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        head(ego, agents, bev, noise=noise)
    e1.record(); torch.cuda.synchronize()
    head.set_profiling(True); head(ego, agents, bev, noise=noise); prof = head.stage_profile(); head.set_profiling(False)
    print(json.dumps({"gran": gran, "seg": seg, "ms": e0.elapsed_time(e1) / 10, "bev_layout": prof["bev_layout"]["ms"], "conv": prof["conv"]["ms"], "chain": prof["gemm_chain"]["ms"]}))
