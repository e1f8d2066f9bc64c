"""Timing of the non-headline BASELINE configs (GPU box): fp32 B=256, stress shape."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth

def run(tag, precision, B, layers, anchors, steps, hw):
    sd = synth.make_state_dict(num_layers=layers, num_anchors=anchors)
    cfg = HeadConfig(num_decoder_layers=layers, step_num=steps)
    head = TrajectoryHead(8, 1024, 256, None, cfg, plan_anchor=sd["plan_anchor"].numpy(), precision=precision)
    head.load_state_dict(sd); head = head.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(7)
    ego = torch.randn(B, 1, 256, device="cuda", generator=g); agents = torch.randn(B, 30, 256, device="cuda", generator=g)
    bev = torch.randn(B, 256, hw, hw, device="cuda", generator=g); noise = torch.randn(B, anchors, 8, 2, device="cuda", generator=g)
    for _ in range(3):
        out = head(ego, agents, bev, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        out = head(ego, agents, bev, noise=noise)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nu = head.debug_tap("nuniq", __import__("numpy").int32)[:B]
    print(f"{tag}: B={B} {precision} {ms:.3f} ms/forward, {B / ms * 1e3:.0f} scenes/s, launches {head.last_launch_count()}, "
          f"unique px/scene-call mean {nu.mean():.0f} max {nu.max()}, finite {bool(torch.isfinite(out['trajectory']).all())}", flush=True)

run("default fp32 (configs[1])", "fp32", 256, 2, 20, 2, 64)
run("default bf16 B=256", "bf16", 256, 2, 20, 2, 64)
run("stress bf16 (configs[4])", "bf16", 512, 4, 64, 3, 128)
run("stress fp32", "fp32", 32, 4, 64, 3, 128)
