"""GPU diagnostic of the scene-tile chain engine: fused vs per-Linear launches vs golden."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth  # noqa: E402


def make(chain, layers=2, anchors=20, steps=2):
    sd = synth.make_state_dict(num_layers=layers, num_anchors=anchors)
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(num_decoder_layers=layers, step_num=steps),
                          plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
    head.load_state_dict(sd)
    head = head.cuda().eval()
    head.set_option("chain_engine", chain)
    return head


def run(head, B, anchors=20, hw=(64, 64)):
    ft = synth.make_features(B, bev_h=hw[0], bev_w=hw[1])
    nz = synth.make_noise(B, num_anchors=anchors)
    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(), noise=nz.cuda())
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def cmp(tag, a, b):
    d = {k: float(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max()) for k in ("trajectory_modes", "trajectory_scores")}
    d["mode_agree"] = float((a["mode_idx"] == b["mode_idx"]).mean())
    d["finite"] = bool(np.isfinite(a["trajectory_modes"]).all())
    print(tag, json.dumps(d), flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if what == "parity":
        for B in (30, 64, 256):
            f = run(make(1), B)
            u = run(make(0), B)
            cmp(f"fused_vs_unfused_B{B}", f, u)
            z = np.load(os.path.join(ROOT, "tests", "golden", "default_b256.npz"))
            ref = {k: z[k][:B] for k in ("trajectory_modes", "trajectory_scores", "mode_idx")}
            cmp(f"fused_vs_golden_B{B}", f, ref)
            cmp(f"unfused_vs_golden_B{B}", u, ref)
        f = run(make(1, 4, 64, 3), 2, 64, (128, 128))
        z = np.load(os.path.join(ROOT, "tests", "golden", "stress_b2.npz"))
        cmp("fused_stress_vs_golden", f, {k: z[k] for k in ("trajectory_modes", "trajectory_scores", "mode_idx")})
    elif what == "time":
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
        seg = int(sys.argv[3]) if len(sys.argv) > 3 else 8
        pconv = int(sys.argv[4]) if len(sys.argv) > 4 else 1
        for chain in (1,) if len(sys.argv) > 3 else (1, 0):
            head = make(chain)
            head.set_option("layout_segment", seg)
            head.set_option("persistent_conv", pconv)
            g = torch.Generator(device="cuda").manual_seed(3000)
            ego = torch.randn(B, 1, 256, device="cuda", generator=g)
            agents = torch.randn(B, 30, 256, device="cuda", generator=g)
            bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
            noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
            for _ in range(3):
                head(ego, agents, bev, noise=noise)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                head(ego, agents, bev, noise=noise)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            head.set_profiling(True)
            head(ego, agents, bev, noise=noise)
            prof = head.stage_profile()
            head.set_profiling(False)
            print(json.dumps({"chain": chain, "B": B, "seg": seg, "pconv": pconv, "ms": ms, "scenes_per_s": B / ms * 1e3,
                              "launches": head.last_launch_count(),
                              "stage_ms": {k: round(v["ms"], 3) for k, v in prof.items()}}), flush=True)
            del head


def timeline(B=4096, launch=1):
    """clock64 stamps of CTA 0's second tile in one chain launch (option chain_timeline)."""
    head = make(1)
    g = torch.Generator(device="cuda").manual_seed(3000)
    ego = torch.randn(B, 1, 256, device="cuda", generator=g)
    agents = torch.randn(B, 30, 256, device="cuda", generator=g)
    bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
    noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
    head(ego, agents, bev, noise=noise)
    head.set_option("chain_timeline", launch)
    head.set_profiling(True)        # (the resident engine is bypassed; B is large anyway)
    head(ego, agents, bev, noise=noise)
    torch.cuda.synchronize()
    d = head.debug_tap("dbg", np.int64)
    head.set_profiling(False)
    t0 = int(d[0])
    names = ["x1", "attn", "ln2ego", "relu_f0", "relu_f1", "relu_f2", "relu_f3", "ln_film", "reg0", "reg2",
             "tail", "cls0_ln", "score"]
    print(f"chain launch {launch}: CTA 0, second tile; cycles relative to the tile start")
    prev_end = 0
    for j in range(13):
        w0, acc, end = (int(d[1 + 4 * j + k]) - t0 for k in range(3))
        m0, m1 = int(d[128 + 2 * j]) - t0, int(d[128 + 2 * j + 1]) - t0
        if end < 0 or end > 10_000_000:
            break
        print(f"step {j:2d} {names[j]:8s} mma: operands ready {m0:7d}, issued {m1:7d} (issue {m1 - m0:6d}) | "
              f"compute: wait from {w0:7d}, acc ready {acc:7d} (waited {acc - w0:6d}), epilogue done {end:7d} "
              f"(epilogue {end - acc:6d})")
        prev_end = end
    print("tile total", prev_end)
    print("attention: q epilogue done", int(d[64]) - t0, "barrier passed", int(d[65]) - t0)
    for sl in range(6):
        a, b, c = (int(d[66 + 3 * sl + k]) - t0 for k in range(3))
        print(f"  scene {sl}: wait K|V from {a} until {b} ({b - a}), compute until {c} ({c - b})")


def conv_timeline(B=4096, launch=1):
    head = make(1)
    g = torch.Generator(device="cuda").manual_seed(3000)
    ego = torch.randn(B, 1, 256, device="cuda", generator=g)
    agents = torch.randn(B, 30, 256, device="cuda", generator=g)
    bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
    noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
    head(ego, agents, bev, noise=noise)
    head.set_option("conv_timeline", launch)
    head.set_profiling(True)
    head(ego, agents, bev, noise=noise)
    torch.cuda.synchronize()
    d = head.debug_tap("dbg", np.int64)[256:]
    nu = head.debug_tap("nuniq", np.int32)
    head.set_profiling(False)
    t0 = int(d[0])
    rel = lambda i: int(d[i]) - t0
    print(f"conv launch {launch}: CTA 0, second scene (scene 148, nuniq {nu[148]}), first pass; cycles since the producers were ready")
    print("producer: passgo ok", rel(1), "| chunk issue starts (every 4th):", [rel(2 + k) for k in range(9)])
    print("mma: operands ready (every 4th chunk):", [rel(16 + k) for k in range(9)], "| committed", rel(25))
    print("epilogue: wait from", rel(32), "accum ok", rel(33), "| half0 drained", rel(34), "synced", rel(35), "combined", rel(36),
          "| half1 drained", rel(37), "synced", rel(38), "combined", rel(39), "| done", rel(40))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "convtl":
        conv_timeline(int(sys.argv[2]) if len(sys.argv) > 2 else 4096, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "timeline":
        timeline(int(sys.argv[2]) if len(sys.argv) > 2 else 4096, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
        sys.exit(0)
    main()
