"""GPU diagnostic: stage-by-stage comparison of the CUDA path with the oracle trace.

Usage (on the GPU box):  python tools/gpu_diag.py [fp32|bf16|gemm32|gemm16|all]
Uses a 1-step / 1-layer head so that the debug taps of the last call correspond to
``s0.l0.*`` of the oracle trace, then the default 2x2 configuration end to end.
Test infrastructure: imports oracle/.
"""
import os
import sys
import time

import numpy as np
import torch

os.environ.setdefault("DDH_DEBUG_TAPS", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402
from oracle import head_oracle  # noqa: E402


def err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()), float(np.abs(b).max())


def run_gemm(prec):
    import ctypes as C
    lib = _lib.load()
    shp = _lib.Shape(20, 8, 256, 1024, 8, 30, 256, 64, 64, 2, 2, 8, 32.0, 32.0)
    hp = C.c_void_p()
    _lib.check(lib, None, lib.ddh_create(C.byref(shp), C.byref(hp)), "create")
    for (M, N, K) in ((128, 256, 64), (128, 256, 256), (300, 256, 512), (1000, 1024, 256), (77, 256, 1024)):
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        Cc = torch.full((M, N), float("nan"), device="cuda")
        rc = lib.ddh_test_gemm(hp, A.data_ptr(), W.data_ptr(), b.data_ptr(), Cc.data_ptr(), M, N, K,
                               prec, None)
        _lib.check(lib, hp, rc, "ddh_test_gemm")
        torch.cuda.synchronize()
        if prec == 1:
            ref = (A.bfloat16().double() @ W.bfloat16().double().t() + b.double())
        else:
            ref = A.double() @ W.double().t() + b.double()
        e = (Cc.double() - ref).abs().max().item()
        print(f"  gemm prec={prec} M={M} N={N} K={K}: max err {e:.3e} (ref max {ref.abs().max().item():.2f})"
              f" nan={int(torch.isnan(Cc).sum())}", flush=True)
    lib.ddh_destroy(hp)


def run_stage(precision, B=3):
    cfg = HeadConfig(num_decoder_layers=1, step_num=1)
    sd = synth.make_state_dict(num_layers=1)
    head = TrajectoryHead(8, 1024, 256, None, cfg, plan_anchor=sd["plan_anchor"].numpy(),
                          precision=precision)
    head.load_state_dict(sd)
    head = head.cuda().eval()
    ft = synth.make_features(B)
    nz = synth.make_noise(B)
    trace = {}
    ref = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz,
                                   num_layers=1, step_num=1, trace=trace)
    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(),
               noise=nz.cuda())
    torch.cuda.synchronize()
    A = 20
    print(f"[stage {precision}] launches={head.last_launch_count()}")
    for tap, key, shape in (("q0", "s0.q0", (B, A, 256)), ("x1", "s0.l0.x1", (B, A, 256)),
                            ("x2", "s0.l0.x2", (B, A, 256)), ("x3", "s0.l0.x3", (B, A, 256))):
        got = head.debug_tap(tap).reshape(shape)
        e, m = err(got, trace[key].numpy())
        print(f"  tap {tap:4s}: max err {e:.3e} (ref max {m:.2f})", flush=True)
    nu = head.debug_tap("nuniq", np.int32)
    print("  nuniq:", nu[:B])
    e, m = err(out["trajectory_modes"].cpu().numpy(), ref["trajectory_modes"].numpy())
    print(f"  modes: max err {e:.3e} m (ref max {m:.2f})")
    e, m = err(out["trajectory_scores"].cpu().numpy(), ref["trajectory_scores"].numpy())
    print(f"  scores: max err {e:.3e}")
    print("  mode idx got", out["mode_idx"].cpu().tolist(), "ref", ref["mode_idx"].tolist(), flush=True)


def run_full(precision, B=8):
    sd = synth.make_state_dict()
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(),
                          precision=precision)
    head.load_state_dict(sd)
    head = head.cuda().eval()
    ft = synth.make_features(B)
    nz = synth.make_noise(B)
    ref = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(),
               noise=nz.cuda())
    torch.cuda.synchronize()
    e, m = err(out["trajectory_modes"].cpu().numpy()[..., :2], ref["trajectory_modes"].numpy()[..., :2])
    eh, _ = err(out["trajectory_modes"].cpu().numpy()[..., 2], ref["trajectory_modes"].numpy()[..., 2])
    es, _ = err(out["trajectory_scores"].cpu().numpy(), ref["trajectory_scores"].numpy())
    agree = (out["mode_idx"].cpu() == ref["mode_idx"]).float().mean().item()
    print(f"[full {precision}] B={B} launches={head.last_launch_count()} xy err {e:.3e} m, heading err {eh:.3e} rad, "
          f"score err {es:.3e}, mode agreement {agree:.3f}", flush=True)
    # timing
    ins = [ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda()]
    nzc = nz.cuda()
    for _ in range(3):
        head(*ins, noise=nzc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        head(*ins, noise=nzc)
    torch.cuda.synchronize()
    print(f"  {1e3 * (time.perf_counter() - t0) / 10:.3f} ms per forward of {B} scenes", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), _lib.load().ddh_build_info().decode(), flush=True)
    if what in ("gemm32", "all"):
        run_gemm(0)
    if what in ("fp32", "all"):
        run_stage("fp32")
        run_full("fp32")
    if what in ("gemm16", "all"):
        run_gemm(1)
    if what in ("bf16", "all"):
        run_stage("bf16")
        run_full("bf16")
