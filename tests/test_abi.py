"""CPU: the C-ABI library loads, exports every symbol include/ddh.h declares, validates shapes,
restates the DDIM table, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from diffusiondrive_b200 import _lib, ddim_alphas_cumprod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from diffusiondrive_b200 import build
        build.build()
    return _lib.load()


def _default_shape(**kw):
    d = dict(num_anchors=20, num_poses=8, d_model=256, d_ffn=1024, num_heads=8, num_agents=30,
             bev_channels=256, bev_h=64, bev_w=64, num_layers=2, num_steps=2, trunc_timestep=8,
             lidar_max_x=32.0, lidar_max_y=32.0)
    d.update(kw)
    return _lib.Shape(**d)


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ddh.h")).read()
    declared = set(re.findall(r"DDH_API[^;(]*?\b(ddh_\w+)\s*\(", hdr))
    assert len(declared) >= 15
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ddh_abi_version() == 1
    assert b"sm_100a" in lib.ddh_build_info()


def test_struct_layouts():
    assert C.sizeof(_lib.Shape) == 14 * 4
    assert C.sizeof(_lib.LayerWeights) == 42 * 8
    assert C.sizeof(_lib.WeightPtrs) == 12 * 8


def test_create_validates_shape(lib):
    h = C.c_void_p()
    assert lib.ddh_create(C.byref(_default_shape()), C.byref(h)) == 0
    lib.ddh_destroy(h)
    for bad in (dict(d_model=128), dict(num_poses=6), dict(num_agents=33), dict(num_agents=0),
                dict(bev_h=300, bev_w=300), dict(num_steps=0), dict(d_ffn=1000),
                dict(bev_channels=64), dict(lidar_max_x=0.0)):
        h = C.c_void_p()
        rc = lib.ddh_create(C.byref(_default_shape(**bad)), C.byref(h))
        assert rc == -2, bad
        assert b"unsupported" in lib.ddh_last_error(None)
        assert not h.value
    assert lib.ddh_create(None, C.byref(h)) == -1


def test_default_ddim_table_matches_torch(lib, golden_dir):
    h = C.c_void_p()
    assert lib.ddh_create(C.byref(_default_shape()), C.byref(h)) == 0
    buf = (C.c_float * 1000)()
    assert lib.ddh_get_alphas_cumprod(h, buf, 1000) == 0
    got = np.frombuffer(buf, dtype=np.float32)
    want = ddim_alphas_cumprod().numpy()
    # the C restatement of torch.linspace/cumprod agrees to a few ulp over all 1000 steps
    assert np.abs(got / want - 1).max() < 2e-6
    z = np.load(os.path.join(golden_dir, "ddim_alphas_cumprod.npz"))["alphas_cumprod"]
    assert np.array_equal(want[:64], z)
    # override path
    t = torch.linspace(0.9, 0.1, 1000)
    assert lib.ddh_set_alphas_cumprod(h, C.cast(t.data_ptr(), C.POINTER(C.c_float)), 1000) == 0
    assert lib.ddh_get_alphas_cumprod(h, buf, 1000) == 0
    assert np.array_equal(np.frombuffer(buf, dtype=np.float32), t.numpy())
    assert lib.ddh_set_alphas_cumprod(h, C.cast(t.data_ptr(), C.POINTER(C.c_float)), 5) == -1
    lib.ddh_destroy(h)


def test_forward_requires_pack_and_device(lib):
    h = C.c_void_p()
    assert lib.ddh_create(C.byref(_default_shape()), C.byref(h)) == 0
    one = 16
    assert lib.ddh_forward(h, one, one, one, 0, 0, one, None, None, None, None, 1, None) == -3
    assert b"not packed" in lib.ddh_last_error(h)
    assert lib.ddh_reserve(h, 4) == -3
    assert lib.ddh_workspace_bytes(h, 256) > 256 * 4 * 1024 * 1024
    if not torch.cuda.is_available():
        wp = _lib.WeightPtrs()
        lw = (_lib.LayerWeights * 2)()
        wp.layers = C.cast(lw, C.POINTER(_lib.LayerWeights))
        rc = lib.ddh_pack_weights(h, C.byref(wp), 0, None)
        assert rc == -4
        assert b"no CUDA device" in lib.ddh_last_error(h)
    lib.ddh_destroy(h)


def test_torch_binding_is_built_and_declines_cpu_tensors():
    """The optional C++ binding above the C ABI (csrc/torch_binding.cpp): built in-tree by build.py,
    importable without a GPU, and it returns None (general Python path) for anything that is not a
    plain CUDA call -- here CPU tensors -- instead of touching the C ABI."""
    import torch
    from diffusiondrive_b200 import build, trajectory_head
    assert os.path.exists(build.build_torch_binding())
    mod = trajectory_head._torch_binding()
    assert mod is not None and hasattr(mod, "forward_fast")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):   # no CUDA runtime to ask for the current device
            mod.forward_fast(0, 0, torch.zeros(1, 1, 256), torch.zeros(1, 30, 256), torch.zeros(1, 256, 64, 64),
                             None, 0, 20, 8, 30, 256, 64, 64, 0)
