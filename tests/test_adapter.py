"""C++ adapter (base_local_planner generator / critic / scored-sampling interfaces over the C ABI)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "humap_local_planner_b200", "adapter")
BIN = os.path.join(ADAPTER, "_build", "adapter_test")


def _build():
    from humap_local_planner_b200 import build as b
    b.build_library()
    subprocess.run(["make", "-s", "-C", ADAPTER], check=True)


def test_adapter_builds_and_fails_loudly_without_gpu():
    import torch
    _build()
    assert os.path.exists(BIN)
    if torch.cuda.is_available():
        return
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_adapter_scored_sampling_equals_direct_plan():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ADAPTER_TEST_OK" in r.stdout, r.stdout + r.stderr
