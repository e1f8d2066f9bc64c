"""Host-side cost of one batch-1 TrajectoryHead call (the part of the eager latency that is not the
kernel): the whole Python call, the bare ctypes call with pre-allocated outputs, and its pieces.

Usage (GPU box): python tools/host_overhead.py
"""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusiondrive_b200 import HeadConfig, TrajectoryHead, synth, _lib  # noqa: E402


def per_call(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        best = min(best, dt)
    return best


def main():
    sd = synth.make_state_dict()
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision="bf16")
    head.load_state_dict(sd)
    head = head.cuda().eval()
    head.frozen = True
    ft = synth.make_features(1)
    ego, agents, bev = ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda()
    nz = synth.make_noise(1).cuda()
    out = head(ego, agents, bev, noise=nz)
    torch.cuda.synchronize()
    # the calls below queue kernels back to back: the per-call figure is host time per launch as long
    # as it exceeds the kernel time; so also report the kernel time
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        head(ego, agents, bev, noise=nz)
    b.record()
    torch.cuda.synchronize()
    print(f"back-to-back forward: {a.elapsed_time(b) * 10:.1f} us per call (device-limited when above the host figures)")
    print(f"head(...) whole Python call        : {per_call(lambda: head(ego, agents, bev, noise=nz), 500):7.2f} us")
    print(f"head.forward_test(...)             : {per_call(lambda: head.forward_test(ego, agents, bev, noise=nz), 500):7.2f} us")
    lib, h = head._lib, head._handle
    traj, modes, scores, idx = out["trajectory"], out["trajectory_modes"], out["trajectory_scores"], out["mode_idx"]
    stream = torch.cuda.current_stream().cuda_stream
    args = (h, ego.data_ptr(), agents.data_ptr(), bev.data_ptr(), _lib.F32, _lib.NCHW, nz.data_ptr(), traj.data_ptr(),
            modes.data_ptr(), scores.data_ptr(), idx.data_ptr(), 1, stream)
    print(f"lib.ddh_forward(*cached args)      : {per_call(lambda: lib.ddh_forward(*args), 500):7.2f} us")
    print(f"8 x tensor.data_ptr()              : {per_call(lambda: [t.data_ptr() for t in (ego, agents, bev, nz, traj, modes, scores, idx)]):7.2f} us")
    print(f"2 x torch.empty                    : {per_call(lambda: (torch.empty((504,), dtype=torch.float32, device='cuda'), torch.empty((1,), dtype=torch.int64, device='cuda'))):7.2f} us")
    print(f"torch._C._cuda_getCurrentRawStream : {per_call(lambda: torch._C._cuda_getCurrentRawStream(0)):7.2f} us")
    print(f"torch.cuda.current_device()        : {per_call(lambda: torch.cuda.current_device()):7.2f} us")
    print(f"event.record()                     : {per_call(lambda: a.record()):7.2f} us")
    # eager latency as bench.py measures it
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for x, y in evs:
        x.record()
        head(ego, agents, bev, noise=nz)
        y.record()
        y.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in evs)
    print(f"eager event-timed p50 {ts[len(ts) // 2]:.1f} us")
    for x, y in evs:
        x.record()
        lib.ddh_forward(*args)
        y.record()
        y.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in evs)
    print(f"bare ctypes call event-timed p50 {ts[len(ts) // 2]:.1f} us")


if __name__ == "__main__":
    main()
