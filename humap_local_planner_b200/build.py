"""In-tree build of the CUDA shared library (sm_100a only) and of the C++ adapter test binary."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libhmp_planner.so")
LIB_CHECK = os.path.join(LIB_DIR, "libhmp_planner_check.so")   # -DHMP_BOUNDS_CHECK debug build (tests/test_gpu_bounds.py)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force: bool = False, verbose: bool = False, check: bool = False) -> str:
    """check=True builds the bounds-checked debug variant (every scene-derived shared / global index asserted on the device)."""
    srcs = [os.path.join(CSRC, "hmp_kernels.cu"), os.path.join(CSRC, "hmp_api.cu")]
    deps = srcs + [os.path.join(CSRC, "hmp_device.h"), os.path.join(CSRC, "hmp_sweep_tpc.inl"), os.path.join(HERE, "..", "include", "hmp_planner.h")]
    target = LIB_CHECK if check else LIB
    if force or _stale(target, deps):
        os.makedirs(LIB_DIR, exist_ok=True)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-DHMP_BOUNDS_CHECK"] if check else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + srcs
        subprocess.run(cmd, check=True, cwd=CSRC)
    return target
