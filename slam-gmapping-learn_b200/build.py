"""In-tree build of the CUDA library (sm_100a only) and of the C++ host layer.

    python -m build  (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  The outputs (libgm_b200.so, libgridfastslam_b200.so, gfs_b200) stay
in this directory: they are git-ignored but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libgm_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # parity: the reference x86-64 build has no FMA contraction
              "--fmad=false", "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return p


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("gm_kernels.cu", "gm_api.cu", "gm_dist.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("gm_kernels.h", "gm_device.cuh", "gm_ctx.h")] + [os.path.join(ROOT, "include", "gm_b200.h")]
    if force or _newer(LIB, deps):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs + ["-ldl"]
        subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines):
    """an experimental build of the CUDA library with extra -D flags -> variants/libgm_b200.<name>.so (for A/B runs on the GPU box:
    copy it over libgm_b200.so)"""
    srcs = [os.path.join(CSRC, f) for f in ("gm_kernels.cu", "gm_api.cu", "gm_dist.cu")]
    out = os.path.join(HERE, "variants", "libgm_b200.%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([nvcc_path()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + srcs + ["-ldl"])
    return out


def build_host(force=False):
    """C++ host layer (GMapping::GridSlamProcessor over the C ABI) + log-driven CLI, if present."""
    mk = os.path.join(HOST, "Makefile")
    if os.path.exists(mk):
        subprocess.check_call(["make", "-s", "-C", HOST] + (["-B"] if force else []))


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)
    return LIB


if __name__ == "__main__":
    import sys
    if "--variant" in sys.argv:  # python build.py --variant r96 GM_SM_REGS=96
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        build_all(force="-f" in sys.argv, verbose="-v" in sys.argv)
        print(LIB)
