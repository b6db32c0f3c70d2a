"""Shared definition of the reference-generated golden fixtures (tests/golden/ref_*.npz): which cycles, which candidate
indices, and a fingerprint of every input so that a drifted scene generator fails loudly instead of comparing apples
with oranges. Used by tools/make_ref_golden.py (writer) and the parity tests (readers)."""
import ctypes as C
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (config, seed, number of sampled candidates) -- the same sampling rule as tests/test_gpu_parity.py::_cycle
CYCLE_CASES = [("cfg0", 0, 72), ("cfg0", 1, 72), ("cfg0", 2, 72), ("cfg0", 3, 72), ("cfg0", 4, 72), ("cfg0", 5, 72),
               ("cfg1", 0, 256), ("cfg1", 1, 256), ("cfg2", 0, 96), ("cfg2", 1, 64)]
FIS_SEED, FIS_N = 1234, 4000


def sample_indices(n_candidates: int, n_sample: int) -> np.ndarray:
    return np.unique(np.linspace(0, n_candidates - 1, min(n_sample, n_candidates)).astype(np.int32))


def fixture_path(name: str, seed: int, n_sample: int) -> str:
    return os.path.join(ROOT, "tests", "golden", f"ref_cycle_{name}_s{seed}_n{n_sample}.npz")


def _bytes(struct) -> bytes:
    return C.string_at(C.addressof(struct), C.sizeof(struct))


def scene_fingerprint(sc, params, smp) -> bytes:
    """sha256 over everything the planner reads: parameters, sampling, world records, costmap, MapGrids, footprint."""
    h = hashlib.sha256()
    h.update(_bytes(params))
    h.update(_bytes(smp))
    w = sc.world
    h.update(np.array([w.robot_x, w.robot_y, w.robot_yaw, w.vel_x, w.vel_y, w.vel_th, w.goal_local_x, w.goal_local_y,
                       w.goal_local_yaw, w.goal_x, w.goal_y, w.goal_yaw]).tobytes())
    h.update(np.array([w.n_obstacles, w.n_people, w.n_groups], dtype=np.int64).tobytes())
    for arr, n in ((sc._obstacles, w.n_obstacles), (sc._people, w.n_people), (sc._groups, w.n_groups)):
        for i in range(n):
            h.update(_bytes(arr[i]))
    h.update(np.ascontiguousarray(sc.cells).tobytes())
    h.update(np.array([sc.origin_x, sc.origin_y, sc.resolution]).tobytes())
    for g in sc.grids:
        h.update(np.ascontiguousarray(g, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(sc.footprint, dtype=np.float64).tobytes())
    h.update(np.array(sc.hv_prev, dtype=np.float64).tobytes())
    return h.digest()


def fis_inputs() -> np.ndarray:
    """(dir_alpha, dir_beta, rel_loc, dist_angle) tuples: uniform angles plus the region boundaries of the rule base."""
    rng = np.random.default_rng(FIS_SEED)
    x = rng.uniform(-np.pi, np.pi, size=(FIS_N, 4))
    edges = np.deg2rad(np.array([-180, -160, -150, -120, -90, -30, -20, 0, 20, 30, 90, 120, 150, 160, 180], dtype=float))
    k = min(len(edges) * 8, FIS_N)
    x[:k, 2] = np.repeat(edges, 8)[:k]          # relative location exactly on term vertices
    x[k:2 * k, 1] = x[k:2 * k, 0]                # equal directions
    x[2 * k:3 * k, 1] = x[2 * k:3 * k, 0] + np.pi  # opposite directions (unwrapped on purpose)
    return x


def load(name: str, seed: int, n_sample: int):
    path = fixture_path(name, seed, n_sample)
    return np.load(path) if os.path.exists(path) else None
