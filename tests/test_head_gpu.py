"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes) behind
``TrajectoryHead.forward``, against (a) the committed outputs of the live reference head
(tests/golden) and (b) the oracle restatement run on the same seeded inputs.

Tolerances are BASELINE.json's: fp32 <= 1e-4 m per waypoint with the identical selected mode
on >= 99.9 % of scenes; bf16 <= 2e-2 m per waypoint (mode agreement reported overall and for
scenes whose reference top-1/top-2 logit margin exceeds 0.05).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from diffusiondrive_b200 import HeadConfig, TrajectoryHead, _lib, synth

pytestmark = pytest.mark.gpu

TOL_FP32_M = 1e-4
TOL_BF16_M = 2e-2


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files if k != "meta"}


def _make_head(precision, num_layers=2, num_anchors=20, step_num=2):
    sd = synth.make_state_dict(num_layers=num_layers, num_anchors=num_anchors)
    cfg = HeadConfig(num_decoder_layers=num_layers, step_num=step_num)
    head = TrajectoryHead(8, 1024, 256, None, cfg, plan_anchor=sd["plan_anchor"].numpy(),
                          precision=precision)
    head.load_state_dict(sd)
    return head.cuda().eval(), sd


def _run(head, B, num_anchors=20, bev_hw=(64, 64), chunk=None):
    ft = synth.make_features(B, bev_h=bev_hw[0], bev_w=bev_hw[1])
    nz = synth.make_noise(B, num_anchors=num_anchors)
    out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(),
               tuple(bev_hw), ft["status_encoding"].cuda(), noise=nz.cuda())
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, ft, nz


def _margin(scores):
    s = np.sort(scores, axis=1)
    return s[:, -1] - s[:, -2]


def _report(tag, out, ref):
    dxy = np.abs(out["trajectory_modes"][..., :2] - ref["trajectory_modes"][..., :2]).max()
    dh = np.abs(out["trajectory_modes"][..., 2] - ref["trajectory_modes"][..., 2]).max()
    ds = np.abs(out["trajectory_scores"] - ref["trajectory_scores"]).max()
    agree = (out["mode_idx"] == ref["mode_idx"])
    big = _margin(ref["trajectory_scores"]) > 0.05
    rec = {"tag": tag, "max_dxy_m": float(dxy), "max_dheading_rad": float(dh),
           "max_dscore": float(ds), "mode_agreement": float(agree.mean()),
           "mode_agreement_margin_gt_0.05": float(agree[big].mean()) if big.any() else None,
           "scenes": int(agree.size), "scenes_margin_gt_0.05": int(big.sum())}
    print("PARITY", json.dumps(rec))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity.jsonl", "a") as fh:
        fh.write(json.dumps(rec) + "\n")
    return rec


# ------------------------------------------------------------------------------ engines
@pytest.mark.parametrize("prec", [0, 1])
def test_gemm_engine_matches_torch(prec):
    lib = _lib.load()
    shp = _lib.Shape(20, 8, 256, 1024, 8, 30, 256, 64, 64, 2, 2, 8, 32.0, 32.0)
    hp = C.c_void_p()
    _lib.check(lib, None, lib.ddh_create(C.byref(shp), C.byref(hp)), "ddh_create")
    try:
        for (M, N, K) in ((1, 256, 64), (128, 256, 256), (129, 512, 512), (1000, 1024, 256),
                          (77, 256, 1024), (4100, 256, 2304)):
            g = torch.Generator().manual_seed(M * 7 + N + K)
            A = torch.randn(M, K, generator=g).cuda()
            W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
            b = torch.randn(N, generator=g).cuda()
            out = torch.full((M, N), float("nan"), device="cuda")
            rc = lib.ddh_test_gemm(hp, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(),
                                   M, N, K, prec, None)
            _lib.check(lib, hp, rc, "ddh_test_gemm")
            if prec == 1:   # bf16 operands, fp32 accumulate: exact up to summation order
                ref = A.bfloat16().double() @ W.bfloat16().double().t() + b.double()
            else:
                ref = A.double() @ W.double().t() + b.double()
            assert not torch.isnan(out).any()
            assert (out.double() - ref).abs().max().item() < 2e-5, (M, N, K)
        assert lib.ddh_test_gemm(hp, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(),
                                 8, 100, 64, prec, None) == -2
    finally:
        lib.ddh_destroy(hp)


# ------------------------------------------------------------------------------ fp32 parity
def test_fp32_b1_matches_reference_golden(golden_dir):
    head, _ = _make_head("fp32")
    out, _, _ = _run(head, 1)
    ref = _load(golden_dir, "default_b1")
    rec = _report("fp32_b1_vs_reference", out, ref)
    assert out["trajectory"].shape == (1, 8, 3) and out["trajectory"].dtype == np.float32
    assert rec["max_dxy_m"] <= TOL_FP32_M and rec["max_dheading_rad"] <= TOL_FP32_M
    assert np.abs(out["trajectory"] - ref["trajectory"]).max() <= TOL_FP32_M
    assert rec["mode_agreement"] == 1.0


def test_fp32_b256_matches_reference_golden(golden_dir):
    head, _ = _make_head("fp32")
    out, _, _ = _run(head, 256)
    ref = _load(golden_dir, "default_b256")
    rec = _report("fp32_b256_vs_reference", out, ref)
    assert rec["max_dxy_m"] <= TOL_FP32_M and rec["max_dheading_rad"] <= TOL_FP32_M
    assert rec["mode_agreement"] >= 0.999
    same = out["mode_idx"] == ref["mode_idx"]
    assert np.abs(out["trajectory"][same] - ref["trajectory"][same]).max() <= TOL_FP32_M
    # the selected trajectory is exactly the selected mode
    pick = out["trajectory_modes"][np.arange(256), out["mode_idx"]]
    assert np.array_equal(pick, out["trajectory"])
    assert np.array_equal(out["mode_idx"], out["trajectory_scores"].argmax(1))


def test_fp32_matches_oracle_live_with_taps():
    """Same seeded inputs through the oracle on this box; also checks internal stages of a
    1-step / 1-layer head whose debug taps correspond to the oracle trace."""
    from oracle import head_oracle
    head, sd = _make_head("fp32", num_layers=1, step_num=1)
    B = 3
    out, ft, nz = _run(head, B)
    trace = {}
    ref = head_oracle.forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz,
                                   num_layers=1, step_num=1, trace=trace)
    for tap, key in (("q0", "s0.q0"), ("x1", "s0.l0.x1"), ("x2", "s0.l0.x2"), ("x3", "s0.l0.x3")):
        got = head.debug_tap(tap).reshape(B, 20, 256)
        assert np.abs(got - trace[key].numpy()).max() < 1e-4, tap
    assert np.abs(head.debug_tap("pts").reshape(B, 20, 8, 2)
                  - ref["trajectory_modes"].numpy()[..., :2]).max() < 1e-4
    nu = head.debug_tap("nuniq", np.int32)[:B]
    assert ((nu > 100) & (nu <= 640)).all()
    assert np.abs(out["trajectory_modes"] - ref["trajectory_modes"].numpy()).max() <= TOL_FP32_M


# ------------------------------------------------------------------------------ bf16 parity
def test_bf16_b256_within_tolerance(golden_dir):
    head, _ = _make_head("bf16")
    out, _, _ = _run(head, 256)
    ref = _load(golden_dir, "default_b256")
    rec = _report("bf16_b256_vs_reference", out, ref)
    assert rec["max_dxy_m"] <= TOL_BF16_M
    assert rec["max_dheading_rad"] <= TOL_BF16_M
    same = out["mode_idx"] == ref["mode_idx"]
    assert np.abs(out["trajectory"][same] - ref["trajectory"][same]).max() <= TOL_BF16_M
    # near-tie scenes may flip under bf16 operands; clear ones must not
    assert rec["mode_agreement"] >= 0.975          # measured 0.988 (3 of 256 scenes, all with margin < 0.02)
    assert rec["mode_agreement_margin_gt_0.05"] == 1.0


def test_bf16_mode_agreement_1024_scenes_and_near_tie_replan(golden_dir):
    """Selected-mode agreement with the live reference over 1024 scenes of the 4096-scene fixture
    (tools/mode_agreement.py runs all 4096): plain bf16 flips only near ties; with
    rescore_margin the near-tie scenes are re-planned by the fp32 engine and agreement reaches
    the north-star's 99.9 %."""
    z = np.load(os.path.join(golden_dir, "default_b4096_scores.npz"))
    N, CH = 1024, 256
    head, _ = _make_head("bf16")
    srt = np.sort(z["trajectory_scores"][:N], axis=1)
    margin = srt[:, -1] - srt[:, -2]
    res = {}
    for m in (None, 0.05):
        head.rescore_margin = m
        idx, tr, resc = [], [], 0
        for s0 in range(0, N, CH):
            ft = synth.make_features(CH, start=s0)
            nz = synth.make_noise(CH, start=s0)
            o = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(), noise=nz.cuda())
            idx.append(o["mode_idx"].cpu().numpy()); tr.append(o["trajectory"].cpu().numpy())
            resc += head.last_rescored
        idx, tr = np.concatenate(idx), np.concatenate(tr)
        agree = idx == z["mode_idx"][:N]
        res[m] = (agree, resc)
        rec = {"tag": f"bf16_modes_1024_rescore_{m}", "mode_agreement": float(agree.mean()),
               "rescored": resc, "flips_max_ref_margin": float(margin[~agree].max()) if (~agree).any() else None}
        print("PARITY", json.dumps(rec))
        with open("gpurun_out/parity.jsonl", "a") as fh:
            fh.write(json.dumps(rec) + "\n")
        assert np.abs(tr[agree] - z["trajectory"][:N][agree]).max() <= TOL_BF16_M
    plain, _ = res[None]
    assert plain.mean() >= 0.975 and plain[margin > 0.05].all()
    fixed, resc = res[0.05]
    assert 0 < resc < N // 2
    assert fixed.mean() >= 0.999


def test_small_batch_engine_matches_reference(golden_dir):
    """B <= 24 in bf16 mode runs the group-resident engine (kernels_res2.cu): the whole forward as ONE
    launch on one 16-CTA cluster per scene.  Same tolerance as the tensor path; it is the engine the
    batch-1 latency number is measured on (B = 1 in its dense mode, see the next test)."""
    ref = _load(golden_dir, "default_b256")
    head, _ = _make_head("bf16")
    ondemand, _ = _make_head("bf16")
    ondemand.set_option("dense_conv", 0)
    outs = {}
    for B in (1, 2, 3, 8, 20):
        for h, tag in ((head, ""), (ondemand, "_ondemand")):
            out, _, _ = _run(h, B)
            sub = {k: v[:B] for k, v in ref.items()}
            rec = _report(f"bf16_resident{tag}_b{B}_vs_reference", out, sub)
            assert rec["max_dxy_m"] <= TOL_BF16_M and rec["max_dheading_rad"] <= TOL_BF16_M
            if B <= 8:
                assert rec["mode_agreement"] == 1.0
            else:   # near-tie scenes may flip under bf16 operands; clear ones must not
                assert rec["mode_agreement_margin_gt_0.05"] in (1.0, None)
            assert h.last_launch_count() == 1
            if tag:
                outs[B] = out
    # with one conv algorithm for every batch size a scene's plan does not depend on the batch it
    # rides in, nor on the run (fixed summation orders)
    for k in ("trajectory_modes", "trajectory_scores", "trajectory", "mode_idx"):
        assert np.array_equal(outs[8][k][:1], outs[1][k])
        assert np.array_equal(outs[8][k][:3], outs[3][k])
        assert np.array_equal(outs[20][k][:8], outs[8][k])
    again, _, _ = _run(ondemand, 8)
    for k in outs[8]:
        assert np.array_equal(again[k], outs[8][k])


def test_resident_engine_dense_mode(golden_dir):
    """Dense mode of the resident engine (default for B = 1, option dense_conv up to 2): helper clusters
    of the same launch run value_proj + ReLU over the WHOLE map (TMA-fed tcgen05 implicit GEMM) and the
    scene clusters gather bilinear corners from it.  Checked: the map itself against torch's conv2d on
    the same bf16 operands, parity with the live-reference golden, layouts, batch independence,
    determinism and that the control words are clean after every launch (repeated calls)."""
    ref = _load(golden_dir, "default_b256")
    head, sd = _make_head("bf16")
    head.set_option("dense_conv", 2)
    head.set_option("debug_taps", 1)
    outs = {}
    for B in (1, 2):
        ft = synth.make_features(B)
        bev = ft["bev_feature"].cuda()
        for rep in range(3):
            out, _, _ = _run(head, B)
            assert head.last_launch_count() == 1
            if rep:
                for k in out:
                    assert np.array_equal(out[k], outs[B][k]), (B, rep, k)
            outs[B] = out
        v = torch.from_numpy(head.debug_tap("dense_v", np.uint16).astype(np.int32)).cuda()
        v = (v << 16).view(torch.float32).view(B, 2, 64, 64, 256)
        for l in range(2):
            w = sd[f"diff_decoder.layers.{l}.cross_bev_attention.value_proj.0.weight"].cuda()
            b = sd[f"diff_decoder.layers.{l}.cross_bev_attention.value_proj.0.bias"].cuda()
            want = torch.nn.functional.conv2d(bev.bfloat16().float(), w.bfloat16().float(), b, padding=1).relu()
            want = want.permute(0, 2, 3, 1)
            # fp32 accumulation in another order, then one bf16 rounding: half an ulp of the largest value
            assert (v[:, l] - want).abs().max().item() <= 2.0 ** -7 * max(1.0, want.max().item()) / 2 + 1e-3
        rec = _report(f"bf16_resident_dense_b{B}_vs_reference", out, {k: x[:B] for k, x in ref.items()})
        assert rec["max_dxy_m"] <= TOL_BF16_M and rec["max_dheading_rad"] <= TOL_BF16_M
        assert rec["mode_agreement"] == 1.0
    for k in outs[1]:
        assert np.array_equal(outs[2][k][:1], outs[1][k]), k
    # NHWC bf16 input: no layout jobs, the TMA reads the caller's map
    ft = synth.make_features(2)
    nz = synth.make_noise(2)
    nhwc = ft["bev_feature"].cuda().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    o2 = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), nhwc, noise=nz.cuda(), bev_layout="NHWC")
    torch.cuda.synchronize()
    for k in outs[2]:
        assert np.array_equal(o2[k].cpu().numpy(), outs[2][k]), k
    # the default: dense at B = 1 only
    plain, _ = _make_head("bf16")
    o1, _, _ = _run(plain, 1)
    for k in outs[1]:
        assert np.array_equal(o1[k], outs[1][k]), k
    # other input forms at B = 1: NHWC fp32 (cast pass in front, then the TMA reads the working copy),
    # NCHW bf16 (layout jobs read bf16), CPU tensors (host call: on-demand engine reading the pinned map
    # in place -- another conv algorithm, same tolerance)
    ft1 = synth.make_features(1)
    nz1 = synth.make_noise(1)
    a1 = (ft1["ego_query"].cuda(), ft1["agents_query"].cuda())
    bev1 = ft1["bev_feature"].cuda()
    o_nhwc32 = plain(*a1, bev1.permute(0, 2, 3, 1).contiguous(), noise=nz1.cuda(), bev_layout="NHWC")
    torch.cuda.synchronize()
    for k in outs[1]:
        assert np.array_equal(o_nhwc32[k].cpu().numpy(), outs[1][k]), k
    bev16 = bev1.bfloat16()
    o_b16 = plain(*a1, bev16, noise=nz1.cuda())
    o_b16_ref = plain(*a1, bev16.float(), noise=nz1.cuda())
    torch.cuda.synchronize()
    for k in outs[1]:
        assert torch.equal(o_b16[k], o_b16_ref[k]), k
    o_host = plain(ft1["ego_query"], ft1["agents_query"], ft1["bev_feature"], noise=nz1)
    assert o_host["trajectory"].device.type == "cpu"
    assert np.abs(o_host["trajectory_modes"].numpy() - outs[1]["trajectory_modes"]).max() <= TOL_BF16_M
    assert np.array_equal(o_host["mode_idx"].numpy(), outs[1]["mode_idx"])


def test_dense_mode_concurrent_launches():
    """Two heads (two handles, two control blocks) launching dense batch-1 forwards on two streams at the
    same time: the helper clusters of either launch may or may not be resident together -- jobs are
    claimed from a counter, so any resident helper makes progress -- and the results equal the
    sequential ones."""
    heads = [_make_head("bf16")[0] for _ in range(2)]
    ins = []
    for i in range(2):
        ft = synth.make_features(3)
        ins.append((ft["ego_query"][i:i + 1].cuda(), ft["agents_query"][i:i + 1].cuda(),
                    ft["bev_feature"][i:i + 1].cuda().contiguous(), synth.make_noise(3)[i:i + 1].cuda().contiguous()))
    want = []
    for h, x in zip(heads, ins):
        want.append({k: v.clone() for k, v in h(*x[:3], noise=x[3]).items()})
        assert h.last_launch_count() == 1
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    for it in range(40):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                outs[i].append(heads[i](*ins[i][:3], noise=ins[i][3]))
    torch.cuda.synchronize()
    for i in range(2):
        for o in outs[i]:
            for k in want[i]:
                assert torch.equal(o[k], want[i][k]), (i, k)


def test_engines_agree():
    """The group-resident engine (default for B <= 24), the scene-tile chain engine (default above,
    and for any B when the resident engine is off) and the per-Linear tensor engine compute the same
    head with the same bf16-operand numerics."""
    B = 2
    base, _, _ = _run(_make_head("bf16")[0], B)
    for opts, launches_min, launches_max in (({"resident_engine": 0}, 20, 40),
                                             ({"resident_engine": 0, "chain_engine": 0}, 41, 200)):
        head, _ = _make_head("bf16")
        for k, v in opts.items():
            head.set_option(k, v)
        out, _, _ = _run(head, B)
        assert launches_min <= head.last_launch_count() <= launches_max, (opts, head.last_launch_count())
        assert np.abs(out["trajectory_modes"] - base["trajectory_modes"]).max() <= TOL_BF16_M
        assert np.array_equal(out["mode_idx"], base["mode_idx"])


def test_chain_engine_matches_per_linear_engine(golden_dir):
    """Scene-tile chain engine (kernels_chain.cu) against one tcgen05 GEMM launch per Linear: same
    operands, bf16 agent attention and bf16 regression-tail input in the chain engine only; ragged
    tile (B not a multiple of 6 scenes per tile), determinism, scene independence."""
    ref = _load(golden_dir, "default_b256")
    B = 77
    chain, _ = _make_head("bf16")
    plain, _ = _make_head("bf16")
    plain.set_option("chain_engine", 0)
    a, _, _ = _run(chain, B)
    b, _, _ = _run(plain, B)
    assert plain.last_launch_count() > chain.last_launch_count() > 1
    assert np.abs(a["trajectory_modes"] - b["trajectory_modes"]).max() <= TOL_BF16_M
    assert np.abs(a["trajectory_scores"] - b["trajectory_scores"]).max() <= 0.05
    sub = {k: v[:B] for k, v in ref.items()}
    rec = _report("bf16_chain_b77_vs_reference", a, sub)
    assert rec["max_dxy_m"] <= TOL_BF16_M and rec["max_dheading_rad"] <= TOL_BF16_M
    again, _, _ = _run(chain, B)
    for k in a:
        assert np.array_equal(again[k], a[k]), k
    small, _, _ = _run(chain, 31)
    for k in a:
        assert np.array_equal(small[k], a[k][:31]), k


def test_resident_engine_nhwc_bf16_input_and_fallback(golden_dir):
    """NHWC bf16 feature maps are gathered in place (no layout pass); shapes outside the engine's
    limits (64 anchors) are served by the other engines."""
    B = 2
    head, _ = _make_head("bf16")
    ft = synth.make_features(B)
    nz = synth.make_noise(B)
    args = (ft["ego_query"].cuda(), ft["agents_query"].cuda())
    out_nchw = head(*args, ft["bev_feature"].cuda(), (64, 64), ft["status_encoding"].cuda(), noise=nz.cuda())
    nhwc = ft["bev_feature"].cuda().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    out_nhwc = head(*args, nhwc, (64, 64), ft["status_encoding"].cuda(), noise=nz.cuda(), bev_layout="NHWC")
    torch.cuda.synchronize()
    assert head.last_launch_count() == 1
    # the NCHW path rounds the same fp32 values to bf16 on the fly: identical operands, identical plan
    for k in out_nchw:
        assert torch.equal(out_nchw[k], out_nhwc[k])
    big, _ = _make_head("bf16", num_layers=4, num_anchors=64, step_num=3)
    out, _, _ = _run(big, 1, num_anchors=64, bev_hw=(128, 128))
    assert big.last_launch_count() > 1
    assert out["trajectory_modes"].shape == (1, 64, 8, 3)


# ------------------------------------------------------------------------------ stress shape
@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32_M), ("bf16", TOL_BF16_M)])
def test_stress_config_matches_reference_golden(golden_dir, precision, tol):
    """64 anchors, 3 denoise steps ([13, 7, 0]), 4 decoder layers, 128x128 BEV."""
    head, _ = _make_head(precision, num_layers=4, num_anchors=64, step_num=3)
    out, _, _ = _run(head, 2, num_anchors=64, bev_hw=(128, 128))
    ref = _load(golden_dir, "stress_b2")
    rec = _report(f"{precision}_stress_b2_vs_reference", out, ref)
    assert out["trajectory_modes"].shape == (2, 64, 8, 3)
    assert rec["max_dxy_m"] <= tol and rec["max_dheading_rad"] <= tol
    if precision == "fp32":
        assert rec["mode_agreement"] == 1.0


# ------------------------------------------------------------------------------ interface
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_host_call_and_layouts_agree(precision):
    """CPU tensors in -> ddh_forward_host; NHWC input skips the layout pass; results identical."""
    head, _ = _make_head(precision)
    B = 5
    ft = synth.make_features(B)
    nz = synth.make_noise(B)
    dev = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(),
               noise=nz.cuda())
    n_dev = head.last_launch_count()
    host = head(ft["ego_query"], ft["agents_query"], ft["bev_feature"], (64, 64),
                ft["status_encoding"], noise=nz)
    for k in dev:
        assert host[k].device.type == "cpu"
        assert torch.equal(host[k], dev[k].cpu()), k
    nhwc = ft["bev_feature"].permute(0, 2, 3, 1).contiguous().cuda()
    o2 = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), nhwc, noise=nz.cuda(),
              bev_layout="NHWC")
    for k in dev:
        assert torch.equal(o2[k], dev[k]), k
    if precision == "bf16":
        o3 = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), nhwc.bfloat16(),
                  noise=nz.cuda(), bev_layout="NHWC")
        # no layout passes (the resident engine runs the whole forward, layout included, as one launch)
        assert head.last_launch_count() < n_dev or n_dev == 1
        for k in dev:
            assert torch.equal(o3[k], dev[k]), k


def test_execution_options_do_not_change_results():
    """Scene-chunk concurrency (two streams) and stage profiling are pure scheduling choices;
    on-demand vs eager BEV layout conversion reads the same pixels: results are bit-identical."""
    B = 37
    ft = synth.make_features(B)
    nz = synth.make_noise(B).cuda()
    args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
    for precision in ("fp32", "bf16"):
        head, _ = _make_head(precision)
        base = {k: v.clone() for k, v in head(*args, noise=nz).items()}
        for chunks in (2, 3, 5):
            head.set_concurrency(chunks, 1)
            out = head(*args, noise=nz)
            for k in base:
                assert torch.equal(out[k], base[k]), (precision, chunks, k)
        head.set_concurrency(1, 512)
        head.set_profiling(True)
        out = head(*args, noise=nz)
        prof = head.stage_profile()
        head.set_profiling(False)
        for k in base:
            assert torch.equal(out[k], base[k]), (precision, "profiling", k)
        if precision == "bf16":   # chain engine: the second denoise step reuses the first one's value rows
            assert prof["conv"]["spans"] == 2 and prof["conv_new"]["spans"] == 2 and prof["combine"]["spans"] == 2
        else:
            assert prof["conv"]["spans"] == 4
        assert prof["conv"]["ms"] > 0
        assert sum(v["ms"] for v in prof.values()) > 0
        # 8-pixel segments converted on demand: a small part of the 512 segments of the map
        done = head.debug_tap("done_seg", np.uint32).reshape(-1, 16)[:B]
        segs = np.array([sum(bin(int(x)).count("1") for x in row) for row in done])
        assert (segs > 30).all() and (segs < 256).all()
    for precision in ("fp32", "bf16"):
        eager, _ = _make_head(precision)
        eager.set_option("lazy_layout", 0)
        seg16, _ = _make_head(precision)
        seg16.set_option("layout_segment", 16)
        lazy_ref, _ = _make_head(precision)
        a = eager(*args, noise=nz)
        b = lazy_ref(*args, noise=nz)
        c = seg16(*args, noise=nz)
        for k in a:
            assert torch.equal(a[k], b[k]), (precision, "lazy-vs-eager", k)
            assert torch.equal(a[k], c[k]), (precision, "segment-16-vs-eager", k)
    # the persistent conv deals scenes to its CTAs on demand (global counter): which CTA runs a scene
    # does not change its result; 600 scenes = four per CTA, ragged
    Bq = 600
    ftq = synth.make_features(Bq)
    nzq = synth.make_noise(Bq).cuda()
    argq = (ftq["ego_query"].cuda(), ftq["agents_query"].cuda(), ftq["bev_feature"].cuda())
    dyn, _ = _make_head("bf16")
    a = {k: v.clone() for k, v in dyn(*argq, noise=nzq).items()}
    dyn.set_option("conv_dynamic", 0)
    b = dyn(*argq, noise=nzq)
    for k in a:
        assert torch.equal(a[k], b[k]), ("conv_dynamic", k)


def test_value_row_reuse_across_denoise_steps():
    """value_proj(bev) of a layer (modules/blocks.py:114) does not depend on the denoise step: the chain
    engine keeps the rows it evaluated and later steps evaluate only pixels no earlier step sampled
    (option conv_reuse).  Checked against the fp32 oracle and against the engine without reuse; with
    conv_reuse = 2 (every pixel of the later steps treated as new: the later steps' kernels do all the
    work, several passes per CTA); with a regression head scaled x10 so that the trajectories move by
    metres between calls (partial overlap); and with three denoise steps."""
    from oracle.head_oracle import forward_test
    B = 64
    ft = synth.make_features(B)
    nz = synth.make_noise(B)
    args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
    head, sd = _make_head("bf16")
    a = {k: v.clone() for k, v in head(*args, noise=nz.cuda()).items()}
    rows = head.debug_tap("conv_rows", np.int32)
    assert rows.shape[0] == 4 and (rows[:2] > 100 * B).all()
    assert (rows[2:] < rows[:2] // 10).all(), rows   # new pixels only
    a2 = head(*args, noise=nz.cuda())
    for k in a:
        assert torch.equal(a[k], a2[k]), ("rerun", k)
    head.set_option("conv_reuse", 0)
    b = head(*args, noise=nz.cuda())
    rows_off = head.debug_tap("conv_rows", np.int32)
    assert (rows_off[2:] > 100 * B).all()
    ref = forward_test(sd, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
    for tag, o in (("reuse", a), ("no-reuse", b)):
        rec = _report("bf16 B=64 " + tag + " vs oracle",
                      {k: v.cpu().numpy() for k, v in o.items()}, {k: v.numpy() for k, v in ref.items()})
        assert rec["max_dxy_m"] <= TOL_BF16_M
    assert (a["trajectory_modes"] - b["trajectory_modes"])[..., :2].abs().max().item() <= 1e-2
    # every pixel of the second step evaluated again by the later steps' kernels: ~250 rows per scene
    # and call through the cross-scene conv (two passes per CTA at 300 scenes) + the gather combine
    Bq = 300
    ftq = synth.make_features(Bq)
    nzq = synth.make_noise(Bq).cuda()
    argq = (ftq["ego_query"].cuda(), ftq["agents_query"].cuda(), ftq["bev_feature"].cuda())
    head.set_option("conv_reuse", 1)
    c = {k: v.clone() for k, v in head(*argq, noise=nzq).items()}
    rows1 = head.debug_tap("conv_rows", np.int32)
    assert (rows1[2:] < rows1[:2] // 10).all(), rows1
    head.set_option("conv_reuse", 2)
    d = {k: v.clone() for k, v in head(*argq, noise=nzq).items()}
    rows2 = head.debug_tap("conv_rows", np.int32)
    assert (np.abs(rows2[2:] - rows2[:2]) < rows2[:2] // 20).all() and (rows2[2:] > 148 * 256).all(), rows2
    head.set_option("conv_reuse", 0)
    e = head(*argq, noise=nzq)
    for tag, x, y in (("all-new vs reuse", d, c), ("all-new vs off", d, e), ("reuse vs off", c, e)):
        dd = (x["trajectory_modes"] - y["trajectory_modes"])[..., :2].abs().max().item()
        assert dd <= 1e-2, (tag, dd)
        assert (x["mode_idx"] == y["mode_idx"]).float().mean().item() >= 0.97, tag
    # trajectories that move by metres between calls: scaled regression head, partial overlap
    sd10 = {k: (v * 10.0 if "plan_reg_branch.4" in k else v) for k, v in sd.items()}
    head10, _ = _make_head("bf16")
    head10.load_state_dict(sd10)
    f = {k: v.clone() for k, v in head10(*args, noise=nz.cuda()).items()}
    rows10 = head10.debug_tap("conv_rows", np.int32)
    assert (rows10[2:] > rows[2:]).all() and (rows10[2:] < rows10[:2]).all(), rows10
    head10.set_option("conv_reuse", 0)
    g = head10(*args, noise=nz.cuda())
    ref10 = forward_test(sd10, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
    rec = _report("bf16 B=64 reg head x10, reuse vs oracle", {k: v.cpu().numpy() for k, v in f.items()},
                  {k: v.numpy() for k, v in ref10.items()})
    assert rec["max_dxy_m"] <= 10 * TOL_BF16_M       # the head's rounding errors scale with its output
    assert (f["trajectory_modes"] - g["trajectory_modes"])[..., :2].abs().max().item() <= 5e-2
    # three denoise steps: rows accumulate over two later steps
    head3, sd3 = _make_head("bf16", step_num=3)
    B3 = 40
    ft3 = synth.make_features(B3)
    nz3 = synth.make_noise(B3)
    arg3 = (ft3["ego_query"].cuda(), ft3["agents_query"].cuda(), ft3["bev_feature"].cuda())
    o3 = {k: v.clone() for k, v in head3(*arg3, noise=nz3.cuda()).items()}
    rows3 = head3.debug_tap("conv_rows", np.int32)
    assert rows3.shape[0] == 6 and (rows3[2:] < rows3[:2].min()).all(), rows3
    ref3 = forward_test(sd3, ft3["ego_query"], ft3["agents_query"], ft3["bev_feature"], nz3, step_num=3)
    rec = _report("bf16 3 steps B=40 reuse vs oracle", {k: v.cpu().numpy() for k, v in o3.items()},
                  {k: v.numpy() for k, v in ref3.items()})
    assert rec["max_dxy_m"] <= TOL_BF16_M
    head3.set_option("conv_reuse", 2)
    p3 = head3(*arg3, noise=nz3.cuda())
    assert (p3["trajectory_modes"] - o3["trajectory_modes"])[..., :2].abs().max().item() <= 1e-2


def test_scene_independence_and_determinism_full_size():
    """Size-independent properties at BASELINE's full size (4096 scenes, bf16): a scene's plan
    does not depend on the batch it rides in nor on its position, and reruns are bit-identical."""
    head, _ = _make_head("bf16")
    B, n = 4096, 64
    ft = synth.make_features(n)
    nz = synth.make_noise(n)
    small = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(),
                 noise=nz.cuda())
    g = torch.Generator(device="cuda").manual_seed(synth.SEED_THROUGHPUT)
    ego = torch.randn(B, 1, 256, device="cuda", generator=g)
    agents = torch.randn(B, 30, 256, device="cuda", generator=g)
    bev = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
    noise = torch.randn(B, 20, 8, 2, device="cuda", generator=g)
    pos = torch.arange(n, device="cuda") * 61 + 5          # scatter the known scenes
    ego[pos], agents[pos], bev[pos], noise[pos] = (ft["ego_query"].cuda(), ft["agents_query"].cuda(),
                                                   ft["bev_feature"].cuda(), nz.cuda())
    big = head(ego, agents, bev, noise=noise)
    for k in small:
        assert torch.equal(big[k][pos], small[k]), k
    again = head(ego, agents, bev, noise=noise)
    for k in big:
        assert torch.equal(again[k], big[k]), k
    assert torch.isfinite(big["trajectory"]).all()
    # reversed batch order -> reversed results
    rev = head(ego.flip(0), agents.flip(0), bev.flip(0), noise=noise.flip(0))
    assert torch.equal(rev["trajectory"].flip(0), big["trajectory"])


def test_edge_cases_vs_oracle():
    """Few agents, odd batch, anchors far outside the BEV grid (all-zero sampling), and points
    exactly on the grid border."""
    from oracle import head_oracle
    sd = synth.make_state_dict()
    for scale, na, B in ((1.0, 5, 3), (40.0, 30, 2), (0.0, 1, 1), (-1.0, 32, 2)):
        sd2 = dict(sd)
        sd2["plan_anchor"] = sd["plan_anchor"] * scale
        head = TrajectoryHead(8, 1024, 256, None, HeadConfig(),
                              plan_anchor=sd2["plan_anchor"].numpy())
        head.load_state_dict(sd2)
        head = head.cuda().eval()
        ft = synth.make_features(B, num_agents=na)
        nz = synth.make_noise(B)
        out = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(),
                   noise=nz.cuda())
        ref = head_oracle.forward_test(sd2, ft["ego_query"], ft["agents_query"],
                                       ft["bev_feature"], nz)
        d = (out["trajectory_modes"].cpu() - ref["trajectory_modes"]).abs().max().item()
        assert d <= TOL_FP32_M, (scale, na, B, d)
        assert torch.equal(out["mode_idx"].cpu(), ref["mode_idx"])


def test_bf16_edge_cases_and_precision_switch():
    """bf16 engines (tensor path and small-batch path) on few/many agents, odd batches and
    out-of-grid anchors; switching precision on one module re-packs."""
    from oracle import head_oracle
    sd = synth.make_state_dict()
    for scale, na, B in ((1.0, 5, 1), (1.0, 5, 3), (40.0, 30, 2), (40.0, 30, 4), (1.0, 32, 9), (1.0, 1, 1)):
        sd2 = dict(sd)
        sd2["plan_anchor"] = sd["plan_anchor"] * scale
        head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd2["plan_anchor"].numpy(),
                              precision="bf16")
        head.load_state_dict(sd2)
        head = head.cuda().eval()
        ft = synth.make_features(B, num_agents=na)
        nz = synth.make_noise(B)
        args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
        out = head(*args, noise=nz.cuda())
        ref = head_oracle.forward_test(sd2, ft["ego_query"], ft["agents_query"], ft["bev_feature"], nz)
        d = (out["trajectory_modes"].cpu() - ref["trajectory_modes"]).abs().max().item()
        assert d <= TOL_BF16_M, (scale, na, B, d)
    head.precision = "fp32"
    out32 = head(*args, noise=nz.cuda())
    d = (out32["trajectory_modes"].cpu() - ref["trajectory_modes"]).abs().max().item()
    assert d <= TOL_FP32_M
    head.precision = "bf16"
    again = head(*args, noise=nz.cuda())
    for k in out:
        assert torch.equal(again[k], out[k]), k


def test_reload_state_dict_repacks():
    head, sd = _make_head("fp32")
    ft = synth.make_features(2)
    nz = synth.make_noise(2).cuda()
    args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
    a = head(*args, noise=nz)["trajectory"].clone()
    sd_b = synth.make_state_dict(seed=123)
    head.load_state_dict(sd_b)
    b = head(*args, noise=nz)["trajectory"].clone()
    assert not torch.equal(a, b)
    fresh = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd_b["plan_anchor"].numpy())
    fresh.load_state_dict(sd_b)
    c = fresh.cuda().eval()(*args, noise=nz)["trajectory"]
    assert torch.equal(b, c)
    with torch.no_grad():                       # in-place edit is picked up by the version check
        head.plan_anchor_encoder[3].bias.add_(0.5)
    d = head(*args, noise=nz)["trajectory"]
    assert not torch.equal(b, d)


def test_unseeded_noise_and_input_errors():
    head, _ = _make_head("fp32")
    ft = synth.make_features(2)
    args = (ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda())
    torch.manual_seed(1)
    a = head(*args)["trajectory"]
    torch.manual_seed(1)
    b = head(*args)["trajectory"]
    assert torch.equal(a, b) and torch.isfinite(a).all()
    with pytest.raises(RuntimeError):
        head(ft["ego_query"].cuda(), ft["agents_query"], ft["bev_feature"].cuda())
    with pytest.raises(TypeError):
        head(args[0], args[1], args[2].half())
