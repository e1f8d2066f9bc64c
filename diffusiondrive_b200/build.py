"""Build the ddh CUDA extension in-tree: diffusiondrive_b200/_ddh.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting
.so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_ddh.so")
STAMP = os.path.join(HERE, "_ddh.so.stamp")
# the same sources with -DDDH_CHECKED: bounded mbarrier waits + index asserts (tests/test_checked_build.py)
OUT_CHECKED = os.path.join(HERE, "_ddh_checked.so")
SOURCES = ["ddh_api.cu", "kernels_simt.cu", "kernels_tc.cu", "kernels_chain.cu", "kernels_producer.cu", "kernels_res2.cu"]
HEADERS = ["common.cuh", "kernels.h", "kernels_chain.h", "kernels_res2.h", "tc_ptx.cuh", "geom.cuh", os.path.join("..", "..", "include", "ddh.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared",
    "-Xptxas", "-v",
]


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    out = OUT_CHECKED if checked else OUT
    stamp = out + ".stamp"
    dig = _digest() + ("-checked" if checked else "")
    if not force and os.path.exists(out) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == dig:
                return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = NVCC_FLAGS + (["-DDDH_CHECKED"] if checked else [])
    cmd = [nvcc] + flags + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "_ddh_checked.build.log" if checked else "_ddh.build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building the ddh extension")
    if verbose:
        print(log)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return out


# ---- optional PyTorch binding above the C ABI (csrc/torch_binding.cpp): removes ~10 us of Python from
# the batch-1 call; the head works without it (ctypes path).  Host-only C++ (g++ through
# torch.utils.cpp_extension / ninja), built in-tree so that it travels to the GPU box.
TORCH_SRC = os.path.join(CSRC, "torch_binding.cpp")
TORCH_DIR = os.path.join(HERE, "_torch_build")
TORCH_OUT = os.path.join(TORCH_DIR, "_ddh_torch.so")


def build_torch_binding(force: bool = False) -> str:
    import torch
    with open(TORCH_SRC, "rb") as fh:
        dig = hashlib.sha256(fh.read() + torch.__version__.encode()).hexdigest()
    stamp = TORCH_OUT + ".stamp"
    if not force and os.path.exists(TORCH_OUT) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == dig:
                return TORCH_OUT
    from torch.utils import cpp_extension
    os.makedirs(TORCH_DIR, exist_ok=True)
    cpp_extension.load(name="_ddh_torch", sources=[TORCH_SRC], build_directory=TORCH_DIR,
                       extra_cflags=["-O2"], with_cuda=True, is_python_module=False, verbose=False)
    if not os.path.exists(TORCH_OUT):
        raise RuntimeError("torch binding did not build")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return TORCH_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build(force="--force" in sys.argv, checked=True))
    print(build_torch_binding(force="--force" in sys.argv))
