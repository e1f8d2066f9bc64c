"""ctypes binding of the ddh C ABI (include/ddh.h).

The shared library is built in-tree by ``diffusiondrive_b200/build.py`` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing, or no CUDA device is present when a
compute entry point is called, the caller gets a ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ddh.so")

DDH_OK = 0
PREC_FP32, PREC_BF16 = 0, 1
F32, BF16 = 0, 1
NCHW, NHWC = 0, 1

_fp = C.c_void_p  # device pointers cross the ABI as plain addresses


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_anchors", "num_poses", "d_model", "d_ffn", "num_heads", "num_agents",
        "bev_channels", "bev_h", "bev_w", "num_layers", "num_steps", "trunc_timestep")] + [
        ("lidar_max_x", C.c_float), ("lidar_max_y", C.c_float)]


LAYER_FIELDS = (
    "bev_attw_w", "bev_attw_b", "bev_out_w", "bev_out_b", "bev_conv_w", "bev_conv_b",
    "agent_in_w", "agent_in_b", "agent_out_w", "agent_out_b",
    "ego_in_w", "ego_in_b", "ego_out_w", "ego_out_b",
    "ffn0_w", "ffn0_b", "ffn2_w", "ffn2_b",
    "norm1_w", "norm1_b", "norm2_w", "norm2_b", "norm3_w", "norm3_b",
    "film_w", "film_b",
    "cls0_w", "cls0_b", "cls_ln2_w", "cls_ln2_b", "cls3_w", "cls3_b", "cls_ln5_w", "cls_ln5_b",
    "cls6_w", "cls6_b",
    "reg0_w", "reg0_b", "reg2_w", "reg2_b", "reg4_w", "reg4_b",
)
GLOBAL_FIELDS = (
    "plan_anchor", "enc0_w", "enc0_b", "enc_ln_w", "enc_ln_b", "enc3_w", "enc3_b",
    "time1_w", "time1_b", "time3_w", "time3_b",
)


class LayerWeights(C.Structure):
    _fields_ = [(n, _fp) for n in LAYER_FIELDS]


class WeightPtrs(C.Structure):
    _fields_ = [(n, _fp) for n in GLOBAL_FIELDS] + [("layers", C.POINTER(LayerWeights))]


class QdecShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_queries", "num_keys", "d_model", "d_ffn", "num_heads",
                                         "num_layers")]


QDEC_LAYER_FIELDS = ("self_in_w", "self_in_b", "self_out_w", "self_out_b", "cross_in_w", "cross_in_b",
                     "cross_out_w", "cross_out_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
                     "norm1_w", "norm1_b", "norm2_w", "norm2_b", "norm3_w", "norm3_b")


class QdecLayerWeights(C.Structure):
    _fields_ = [(n, _fp) for n in QDEC_LAYER_FIELDS]


class QdecWeightPtrs(C.Structure):
    _fields_ = [("query_embedding", _fp), ("layers", C.POINTER(QdecLayerWeights))] + [
        (n, _fp) for n in ("states0_w", "states0_b", "states2_w", "states2_b", "label_w", "label_b")]


# name -> (restype, argtypes); every symbol declared in include/ddh.h
SIGNATURES = {
    "ddh_abi_version": (C.c_int, []),
    "ddh_build_info": (C.c_char_p, []),
    "ddh_create": (C.c_int, [C.POINTER(Shape), C.POINTER(C.c_void_p)]),
    "ddh_destroy": (None, [C.c_void_p]),
    "ddh_last_error": (C.c_char_p, [C.c_void_p]),
    "ddh_set_alphas_cumprod": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "ddh_get_alphas_cumprod": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "ddh_pack_weights": (C.c_int, [C.c_void_p, C.POINTER(WeightPtrs), C.c_int, C.c_void_p]),
    "ddh_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "ddh_reserve": (C.c_int, [C.c_void_p, C.c_int]),
    "ddh_forward": (C.c_int, [C.c_void_p, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp,
                              C.c_int, C.c_void_p]),
    "ddh_forward_host": (C.c_int, [C.c_void_p, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _fp,
                                   _fp, C.c_int, C.c_void_p]),
    "ddh_bev_producer_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ddh_bev_producer": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, _fp, C.c_void_p]),
    "ddh_qdec_create": (C.c_int, [C.POINTER(QdecShape), C.POINTER(C.c_void_p)]),
    "ddh_qdec_destroy": (None, [C.c_void_p]),
    "ddh_qdec_last_error": (C.c_char_p, [C.c_void_p]),
    "ddh_qdec_pack_weights": (C.c_int, [C.c_void_p, C.POINTER(QdecWeightPtrs), C.c_int, C.c_void_p]),
    "ddh_qdec_forward": (C.c_int, [C.c_void_p, _fp, _fp, _fp, _fp, C.c_int, C.c_void_p]),
    "ddh_last_launch_count": (C.c_int, [C.c_void_p]),
    "ddh_set_concurrency": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ddh_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ddh_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ddh_get_profile": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "ddh_debug_copy": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "ddh_test_gemm": (C.c_int, [C.c_void_p, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p]),
}

_lib = None


def use_library(path: str) -> None:
    """Test hook: bind another build of the same ABI (the -DDDH_CHECKED library) for every handle
    created from now on.  Existing ``TrajectoryHead`` instances keep the library they loaded."""
    global _lib, LIB_PATH
    LIB_PATH = path
    _lib = None


def load() -> C.CDLL:
    """Load the extension; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"ddh CUDA extension not found at {LIB_PATH}. Build it with "
            "`python -m diffusiondrive_b200.build` (needs nvcc); there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib: C.CDLL, handle, rc: int, what: str) -> None:
    if rc != DDH_OK:
        msg = lib.ddh_last_error(handle)
        raise RuntimeError(f"{what} failed (ddh_status {rc}): {msg.decode() if msg else ''}")
