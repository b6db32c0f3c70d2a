"""Host-side input generation: the numpy MapGrid wave front equals the oracle's restatement of
base_local_planner::MapGrid, the sampling helpers equal the oracle's, and scenes are deterministic."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
from humap_local_planner_b200 import config, scenes


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("local_goal", [False, True])
def test_wavefront_matches_oracle(seed, local_goal):
    cfg = scenes.CONFIGS["cfg2" if seed % 2 else "cfg0"]
    sc = scenes.make_scene(cfg, seed)
    plan = np.array([[s, 0.02 * s] for s in np.arange(0.0, 4.0, 0.1)])
    mine = scenes.mapgrid_wavefront(sc.cells, sc.origin_x, sc.origin_y, sc.resolution, plan, local_goal)
    ref = np.zeros_like(mine)
    L = ob.lib()
    L.orc_mapgrid_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p]
    L.orc_mapgrid_compute.restype = None
    L.orc_mapgrid_compute(sc.cells.ctypes.data, sc.size_x, sc.size_y, sc.origin_x, sc.origin_y, sc.resolution,
                          plan.ctypes.data, plan.shape[0], 1 if local_goal else 0, ref.ctypes.data)
    assert np.array_equal(mine, ref)
    assert (mine == 0).sum() >= 1


@pytest.mark.parametrize("name,count", [("cfg0", 72), ("cfg1", 16384), ("cfg2", 65536), ("cfg3", 4096)])
def test_candidate_counts(name, count):
    smp = scenes.make_sampling(scenes.CONFIGS[name])
    assert config.count_candidates(smp) == count
    assert ob.num_candidates(smp) == count


def test_scene_is_deterministic_and_sized():
    cfg = scenes.CONFIGS["cfg2"]
    a, b = scenes.make_scene(cfg, 5), scenes.make_scene(cfg, 5)
    assert np.array_equal(a.cells, b.cells) and all(np.array_equal(x, y) for x, y in zip(a.grids, b.grids))
    assert a.world.n_people == 50 and a.world.n_groups == 8 and a.world.n_obstacles == 550
    assert a.cells.shape == (200, 200)
    params = scenes.make_params(cfg)
    assert ob.num_steps(params, a.world) == 50
    assert ob.num_steps(scenes.make_params(scenes.CONFIGS["cfg0"]), a.world) == 35
