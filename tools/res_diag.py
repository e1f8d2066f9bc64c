"""GPU diagnostic of the resident (one-launch) engine: stage taps against the oracle trace on a
1-step / 1-layer head, end-to-end parity on the default head for several batch sizes, the old
chain engine beside it, batch-1 latency (stream launches and CUDA-graph replay) and the
in-kernel clock64 timeline.

Usage (GPU box):  python tools/res_diag.py [stage|full|time|all]
Test infrastructure: imports oracle/.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402
from oracle import head_oracle  # noqa: E402


RES_MODE = 2


def err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()), float(np.abs(b).max())


def make_head(cfg, sd, res):
    """res = 2: group-resident engine (default for B <= 24); 0: chain engine for every batch size."""
    head = TrajectoryHead(8, 1024, 256, None, cfg, plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
    head.load_state_dict(sd)
    head = head.cuda().eval()
    head.set_option("debug_taps", 1)
    head.set_option("resident_engine", 1 if res else 0)
    ft = synth.make_features(1)
    head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(), noise=synth.make_noise(1).cuda())
    torch.cuda.synchronize()
    return head


def run_stage(B=2):
    cfg = HeadConfig(num_decoder_layers=1, step_num=1)
    sd = synth.make_state_dict(num_layers=1)
    ft = synth.make_features(B)
    nz = synth.make_noise(B)
    trace = {}
    ref = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz,
                                   num_layers=1, step_num=1, trace=trace)
    head = make_head(cfg, sd, RES_MODE)
    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(), noise=nz.cuda())
    torch.cuda.synchronize()
    A = 20
    print(f"[stage res] B={B} launches={head.last_launch_count()}", flush=True)
    for tap, key in (("res_q0", "s0.q0"), ("res_x1", "s0.l0.x1")):
        got = head.debug_tap(tap).reshape(B, A, 256)
        e, m = err(got, trace[key].numpy())
        print(f"  tap {tap:8s}: max err {e:.3e} (ref max {m:.2f}) nan={int(np.isnan(got).sum())}", flush=True)
    reg = trace["s0.l0.reg"].numpy()
    pts = trace["s0.pts"].numpy()
    raw = head.debug_tap("res_regraw").reshape(B, A, 8, 3)
    e, m = err(raw[..., :2], reg[..., :2] - pts)
    print(f"  tap regraw  : max err {e:.3e} (ref max {m:.2f})")
    e, m = err(out["trajectory_modes"].cpu().numpy(), ref["trajectory_modes"].numpy())
    print(f"  modes: max err {e:.3e} m (ref max {m:.2f})")
    e, m = err(out["trajectory_scores"].cpu().numpy(), ref["trajectory_scores"].numpy())
    print(f"  scores: max err {e:.3e}")
    print("  mode idx got", out["mode_idx"].cpu().tolist(), "ref", ref["mode_idx"].tolist(), flush=True)


def run_full():
    sd = synth.make_state_dict()
    cfg = HeadConfig()
    heads = {"res": make_head(cfg, sd, RES_MODE), "chain": make_head(cfg, sd, 0)}
    for B in (1, 2, 3, 8):
        ft = synth.make_features(B)
        nz = synth.make_noise(B)
        ref = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
        for name, head in heads.items():
            for layout in ("NCHW", "NHWC"):
                bev = ft["bev_feature"].cuda()
                if layout == "NHWC":
                    bev = bev.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
                out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), bev, noise=nz.cuda(),
                           bev_layout=layout)
                torch.cuda.synchronize()
                m = out["trajectory_modes"].cpu().numpy()
                e, _ = err(m[..., :2], ref["trajectory_modes"].numpy()[..., :2])
                eh, _ = err(m[..., 2], ref["trajectory_modes"].numpy()[..., 2])
                es, _ = err(out["trajectory_scores"].cpu().numpy(), ref["trajectory_scores"].numpy())
                et, _ = err(out["trajectory"].cpu().numpy(), ref["trajectory"].numpy())
                agree = (out["mode_idx"].cpu() == ref["mode_idx"]).float().mean().item()
                print(f"[full {name} {layout}] B={B} launches={head.last_launch_count()} xy {e:.3e} m, "
                      f"heading {eh:.3e}, score {es:.3e}, traj {et:.3e}, modes agree {agree:.3f}", flush=True)


def run_time():
    sd = synth.make_state_dict()
    cfg = HeadConfig()
    modes = (("res", RES_MODE),) if os.environ.get("ONLY_RES") else (("res", RES_MODE), ("chain", 0))
    for name, res in modes:
        head = make_head(cfg, sd, res)
        for B in ((1,) if os.environ.get("ONLY_RES") else (1, 2, 4, 8)):
            ft = synth.make_features(B)
            nz = synth.make_noise(B).cuda()
            ins = [ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda()]
            for _ in range(5):
                head(*ins, noise=nz)
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
            for a, b in evs:
                a.record()
                head(*ins, noise=nz)
                b.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
            print(f"[time {name}] B={B} launches={head.last_launch_count()} p50 {ts[len(ts)//2]:.1f} us "
                  f"min {ts[0]:.1f} us", flush=True)
            if B == 1 and res == RES_MODE:
                dbg_all = head.debug_tap("dbg", np.int64)[:1000]
                dbg = dbg_all[:900]
                n = int((dbg > 0).sum())
                extra = dbg_all[900:916]
                if (extra > 0).any():
                    lab0 = (dbg[:n] >> 48).tolist()
                    clk0 = (dbg[:n] & ((1 << 48) - 1)).tolist()
                    ref = clk0[lab0.index(112)] if 112 in lab0 else clk0[0]
                    print("  per-warp stamps relative to the first mark 112:", [int(x) - ref if x > 0 else None for x in extra.tolist()])
                if n > 2:
                    lab = (dbg[:n] >> 48).tolist()
                    clk = (dbg[:n] & ((1 << 48) - 1)).tolist()
                    print("  timeline: label:+cycles since previous mark")
                    line = []
                    for i in range(1, n):
                        line.append(f"{lab[i]}:+{clk[i] - clk[i - 1]}")
                        if lab[i] < 100:
                            print("   ", " ".join(line))
                            line = []
                    print("  total cycles", clk[-1] - clk[0])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), _lib.load().ddh_build_info().decode(), flush=True)
    if what in ("stage", "all"):
        run_stage()
    if what in ("full", "all"):
        run_full()
    if what in ("time", "all"):
        run_time()
