"""The C-ABI library loads on a CPU-only box and exports every symbol include/hmp_planner.h declares.
No compute entry point is called here (there is no GPU)."""
import ctypes as C
import os
import re

from humap_local_planner_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "hmp_planner.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(capi.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.hmp_abi_version() == 2


def test_struct_sizes_match_the_c_layout():
    # sizes computed by the C compiler for include/hmp_planner.h (gcc, x86-64); a mismatch means the ctypes mirror drifted
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "hmp_planner.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(HmpParams),sizeof(HmpWorld),sizeof(HmpObstacle),sizeof(HmpPerson),sizeof(HmpGroup),' \
          'sizeof(HmpSampling),sizeof(HmpSample),sizeof(HmpResult),sizeof(HmpCosts),sizeof(HmpEnvParams),sizeof(HmpShape),' \
          'sizeof(HmpEquisampled));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")], check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()]
    mine = [C.sizeof(t) for t in (capi.HmpParams, capi.HmpWorld, capi.HmpObstacle, capi.HmpPerson, capi.HmpGroup,
                                  capi.HmpSampling, capi.HmpSample, capi.HmpResult, capi.HmpCosts, capi.HmpEnvParams, capi.HmpShape,
                                  capi.HmpEquisampled)]
    assert mine == sizes


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        return
    lib = capi.load_library()
    ctx = lib.hmp_create(0)
    assert not ctx
    assert b"no CPU fallback" in lib.hmp_last_error() or b"CUDA" in lib.hmp_last_error()
    try:
        capi.Planner(0)
        raise AssertionError("Planner() must raise without a GPU")
    except capi.HmpError as e:
        assert e.code == capi.HMP_E_CUDA
