"""GPU diagnostic of the resident engine's dense mode (whole-map value_proj on helper clusters of the
same launch, kernels_res2.h DenseArgs): the V map against torch's conv2d on the same bf16 operands,
dense vs on-demand outputs vs the oracle for B = 1, 2 and both layouts, latency and the in-kernel
timeline.

Usage (GPU box):  [CHECKED=1] python tools/dense_diag.py [v|full|time|all]
Test infrastructure: imports oracle/.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402
from oracle import head_oracle  # noqa: E402

if os.environ.get("CHECKED"):
    _lib.use_library(os.path.join(ROOT, "diffusiondrive_b200", "_ddh_checked.so"))


def make_head(sd, dense, taps=True):
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
    head.load_state_dict(sd)
    head = head.cuda().eval()
    if taps:
        head.set_option("debug_taps", 1)
    head.set_option("dense_conv", dense)
    return head


def run_v():
    sd = synth.make_state_dict()
    for B in (1, 2):
        head = make_head(sd, 2)
        ft = synth.make_features(B)
        nz = synth.make_noise(B)
        bev = ft["bev_feature"].cuda()
        for rep in range(2):   # twice: the control words must be back to zero after a launch
            head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), bev, noise=nz.cuda())
            torch.cuda.synchronize()
            v = torch.from_numpy(head.debug_tap("dense_v", np.uint16).astype(np.int32)).cuda()
            v = (v << 16).view(torch.float32).view(B, 2, 64, 64, 256)
            for l in range(2):
                w = sd[f"diff_decoder.layers.{l}.cross_bev_attention.value_proj.0.weight"].cuda()
                b = sd[f"diff_decoder.layers.{l}.cross_bev_attention.value_proj.0.bias"].cuda()
                ref = torch.nn.functional.conv2d(bev.bfloat16().float(), w.bfloat16().float(), b, padding=1).relu()
                ref = ref.permute(0, 2, 3, 1)
                d = (v[:, l] - ref).abs()
                print(f"[v] B={B} rep={rep} layer {l}: max |dV| {d.max().item():.3e} (ref max {ref.max().item():.2f}), "
                      f"mean {d.mean().item():.3e}, nonzero frac {float((v[:, l] != 0).float().mean()):.3f}", flush=True)


def run_full():
    sd = synth.make_state_dict()
    heads = {"dense": make_head(sd, 2, taps=False), "ondemand": make_head(sd, 0, taps=False)}
    for B in (1, 2, 3):
        ft = synth.make_features(B)
        nz = synth.make_noise(B)
        ref = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
        outs = {}
        for name, head in heads.items():
            for layout in ("NCHW", "NHWC"):
                bev = ft["bev_feature"].cuda()
                if layout == "NHWC":
                    bev = bev.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
                for rep in range(2):
                    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), bev, noise=nz.cuda(), bev_layout=layout)
                    torch.cuda.synchronize()
                m = out["trajectory_modes"].cpu().numpy()
                r = ref["trajectory_modes"].numpy()
                e = np.abs(m[..., :2] - r[..., :2]).max()
                eh = np.abs(m[..., 2] - r[..., 2]).max()
                es = np.abs(out["trajectory_scores"].cpu().numpy() - ref["trajectory_scores"].numpy()).max()
                agree = (out["mode_idx"].cpu() == ref["mode_idx"]).float().mean().item()
                outs[(name, layout)] = m
                print(f"[full {name} {layout}] B={B} launches={head.last_launch_count()} xy {e:.3e} m, heading {eh:.3e}, "
                      f"score {es:.3e}, modes agree {agree:.3f}", flush=True)
        print(f"   dense NCHW == dense NHWC: {np.array_equal(outs[('dense', 'NCHW')], outs[('dense', 'NHWC')])}; "
              f"dense vs on-demand max {np.abs(outs[('dense', 'NCHW')] - outs[('ondemand', 'NCHW')]).max():.3e}", flush=True)


def run_time():
    sd = synth.make_state_dict()
    for dense in (1, 0):
        for taps in (False, True):
            head = make_head(sd, dense, taps=taps)
            head.frozen = True
            B = 1
            ft = synth.make_features(B)
            nz = synth.make_noise(B).cuda()
            ins = [ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda()]
            for _ in range(5):
                head(*ins, noise=nz)
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
            for a, b in evs:
                a.record()
                head(*ins, noise=nz)
                b.record()
                b.synchronize()
            ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
            line = f"[time dense={dense} taps={int(taps)}] B=1 eager p50 {ts[len(ts)//2]:.1f} us min {ts[0]:.1f} us"
            if not taps:
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        head(*ins, noise=nz)
                torch.cuda.synchronize()
                for _ in range(5):
                    graph.replay()
                torch.cuda.synchronize()
                gd = []
                for a, b in evs:
                    a.record()
                    graph.replay()
                    b.record()
                    b.synchronize()
                    gd.append(a.elapsed_time(b) * 1e3)
                gd.sort()
                line += f"; graph replay p50 {gd[len(gd)//2]:.1f} us min {gd[0]:.1f} us"
            print(line, flush=True)
            if taps and dense:
                hd = head.debug_tap("dbg", np.int64)[920:940]
                rel = [int(x - hd[0]) if x > 0 else None for x in hd[:9]]
                print("  helper CTA of conv job 0 (cycles since role entry): claimed %s, rows ready %s, first stage full %s, "
                      "MMAs issued %s, accumulators ready %s, epilogue done %s (staged %s, lines written %s)" % tuple(rel[1:9]))
                print(f"  a helper CTA's first layout job (job {int(hd[13])}): starts {int(hd[14])}, stored {int(hd[15])}, flagged {int(hd[16])} cycles after role entry")
                print(f"  globaltimer: helper job-0 done {int(hd[10] - hd[11])} ns after the scene cluster's start; "
                      f"scene cluster at its first gather after {int(hd[12] - hd[11])} ns", flush=True)
            if taps:
                dbg = head.debug_tap("dbg", np.int64)[:900]
                n = int((dbg > 0).sum())
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
                    print("  total cycles", clk[-1] - clk[0], flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), _lib.load().ddh_build_info().decode(), flush=True)
    if what in ("v", "all"):
        run_v()
    if what in ("full", "all"):
        run_full()
    if what in ("time", "all"):
        run_time()
