"""B200-native trajectory sampling + scoring hot path of humap_local_planner.

Layout:
  csrc/      CUDA kernels (hmp_kernels.cu), C-ABI host code (hmp_api.cu), device layout (hmp_device.h)
  adapter/   C++ adapter that keeps the base_local_planner generator / critic / scored-sampling API
  capi.py    ctypes binding of include/hmp_planner.h
  config.py  default parameter sets (cfg/HumapPlanner.cfg values flattened into HmpParams)
  scenes.py  seeded synthetic planning cycles for the BASELINE configurations
  sharding.py  scene -> GPU partitioning for batched scenes (no collective on the data path)
"""
from . import capi, config, scenes  # noqa: F401
from .capi import Planner, HmpError, load_library  # noqa: F401

__all__ = ["capi", "config", "scenes", "Planner", "HmpError", "load_library"]
