"""The -DDDH_CHECKED twin of the extension (bounded mbarrier waits that trap instead of hanging,
DDH_ASSERT on the index arithmetic of the pipelines): same ABI, same results, no assert fires on the
engines' main paths.  compute-sanitizer is not available on the GPU pool; this is what stands in."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "diffusiondrive_b200", "_ddh_checked.so")

SCRIPT = r'''
import sys, json
import numpy as np, torch
sys.path.insert(0, %(root)r)
from diffusiondrive_b200 import _lib, HeadConfig, TrajectoryHead, synth
if %(checked)d:
    _lib.use_library(%(lib)r)
info = _lib.load().ddh_build_info().decode()
assert ("checked" in info) == bool(%(checked)d), info
out = {}
for tag, prec, B, opts in (("chain", "bf16", 77, {}), ("resident", "bf16", 3, {}), ("resident_dense", "bf16", 2, {"dense_conv": 2}), ("per_linear", "bf16", 9, {"chain_engine": 0, "resident_engine": 0}),
                           ("conv2", "bf16", 40, {"persistent_conv": 1}), ("fp32", "fp32", 5, {})):
    sd = synth.make_state_dict()
    head = TrajectoryHead(8, 1024, 256, None, HeadConfig(), plan_anchor=sd["plan_anchor"].numpy(), precision=prec)
    head.load_state_dict(sd)
    head = head.cuda().eval()
    for k, v in opts.items():
        head.set_option(k, v)
    ft = synth.make_features(B)
    o = head(ft["ego_query"].cuda(), ft["agents_query"].cuda(), ft["bev_feature"].cuda(), noise=synth.make_noise(B).cuda())
    torch.cuda.synchronize()
    out[tag] = float(o["trajectory_modes"].double().abs().sum()) + float(o["trajectory_scores"].double().sum())
print("RESULT", json.dumps(out))
'''


def _run(checked):
    code = SCRIPT % {"root": ROOT, "lib": CHECKED, "checked": int(checked)}
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DDH_ASSERT" not in res.stdout and "timed out" not in res.stdout
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT")][-1]
    return line


def test_checked_library_is_built_and_exports_the_abi():
    import ctypes
    from diffusiondrive_b200 import _lib
    assert os.path.exists(CHECKED), "python -m diffusiondrive_b200.build builds it"
    lib = ctypes.CDLL(CHECKED)
    for name in _lib.SIGNATURES:
        assert hasattr(lib, name), name
    lib.ddh_build_info.restype = ctypes.c_char_p
    assert b"checked" in lib.ddh_build_info()


@pytest.mark.gpu
def test_checked_build_runs_clean_and_matches():
    """Every engine under the checked build: no assert, no timeout, checksums identical to the
    production build (the checks do not touch the arithmetic)."""
    assert _run(True) == _run(False)
