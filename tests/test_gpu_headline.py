"""The headline configurations against the COMPILED REFERENCE's full-grid outputs (tests/golden/ref_full_*.npz, frozen by
tools/make_ref_golden_full.py from oracle/_ref = the reference's own sources): BASELINE config 3 (cfg2: 65 536 candidates x
50 people x 8 groups x 500 obstacle points, seeds 0-2) and config 4 (cfg3: batched worlds x 4096 candidates, 8 scenes).

north_star gate: "the chosen candidate is identical unless the reference's top-2 totals are within 1e-4 relative"
(reference: findBestTrajectory at src/humap_planner.cpp:1367). Checked for the default precision mode 2 (FP32 sweep + FP64
refinement) in BOTH sweep layouts, for the exact FP64 mode 1, and for hmp_plan_batch.
"""
import os

import numpy as np
import pytest

import golden_cases as gc
from humap_local_planner_b200 import scenes

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _full(name, seed):
    path = os.path.join(GOLD, f"ref_full_{name}_s{seed}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} missing: run tools/make_ref_golden_full.py where /root/reference exists")
    return np.load(path)


def _setup(planner, name, seed, mode):
    planner.set_precision(mode)
    cfg = scenes.CONFIGS[name]
    sc = scenes.make_scene(cfg, seed)
    params = scenes.make_params(cfg)
    smp = scenes.make_sampling(cfg)
    planner.set_params(params)
    planner.set_scene(sc)
    return cfg, sc, params, smp


def _top2_close(ref_totals):
    v = np.sort(ref_totals[ref_totals >= 0])
    return len(v) > 1 and (v[1] - v[0]) <= 1e-4 * abs(v[0])


def _check_winner(res, poses, g, tol):
    ref_best = int(g["best_index"])
    assert res.status == 0
    assert res.best_index == ref_best or _top2_close(g["totals"]), (res.best_index, ref_best, res.best_total, float(g["best_total"]))
    if res.best_index != ref_best:
        return
    assert abs(res.best_total - float(g["best_total"])) <= tol * abs(float(g["best_total"]))
    assert np.allclose(np.array(res.costs), g["best_costs"], rtol=tol, atol=tol, equal_nan=True)
    assert np.abs(np.array([res.xv, res.yv, res.thetav]) - g["best_seed"]).max() <= tol
    if poses is not None:
        bp = g["best_poses"]
        assert poses.shape == bp.shape
        assert np.abs(poses[:, :2] - bp[:, :2]).max() <= tol
        assert np.abs((poses[:, 2] - bp[:, 2] + np.pi) % (2 * np.pi) - np.pi).max() <= tol


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("lay", [1, 2], ids=["warp", "thread"])
def test_cfg2_default_mode_selects_the_references_winner(planner, seed, lay):
    """Mode 2 (what bench.py times): winner == the reference's over the FULL 65 536-candidate grid, winner record (total, 14
    raw critics, seed twist, 50 poses) within 1e-6; the FP32 explored totals of all candidates against the reference's."""
    g = _full("cfg2", seed)
    cfg, sc, params, smp = _setup(planner, "cfg2", seed, 2)
    assert gc.scene_fingerprint(sc, params, smp) == g["fingerprint"].tobytes(), "scene drifted from the fixture"
    planner.set_sweep_layout(lay)
    try:
        res, poses = planner.plan(sc.world, smp)
        assert (planner.last_sweep_mode() != 0) == (lay == 2)
        n_lead = planner.last_num_leaders()
        t_ref = planner.explored_totals(res.n_candidates)     # FP32 totals, leaders replaced by their FP64 totals
        planner.set_precision(0)
        res0, _ = planner.plan(sc.world, smp, want_poses=False)
        t32 = planner.explored_totals(res0.n_candidates)      # pure FP32 totals
    finally:
        planner.set_sweep_layout(0)
        planner.set_precision(0)
    ref = g["totals"]
    assert res.n_candidates == int(g["C"]) == 65536
    _check_winner(res, poses, g, 1e-6)
    # where the reference's winner ranks in the FP32 ordering (the refinement must reach it): reported, and bounded
    ref_best = int(g["best_index"])
    v32 = t32 >= 0
    rank = int((t32[v32] < t32[ref_best]).sum()) if t32[ref_best] >= 0 else -1
    gap = (t32[ref_best] - t32[v32].min()) / t32[v32].min()
    print(f"cfg2 seed {seed} layout {lay}: reference winner {ref_best} has FP32 rank {rank} (+{gap:.2e} rel above the FP32 best), "
          f"{n_lead} leaders refined")
    assert 0 <= rank < n_lead
    # validity and error codes of every candidate
    neg = (t32 < 0) | (ref < 0)
    mism = neg & (t32 != ref)
    # measured (r02a): 1-2 of 65 536 (a footprint vertex or a feasibility test on the edge)
    # measured (r02a/b): seed 0: 1-2 of 65 536; seeds 1, 2 (NO_INFORMATION border, more chaotic rollouts): 9-10 in the thread layout,
    # 62-109 in the warp layout (other summation order of the forces)
    assert mism.sum() <= 160, f"{int(mism.sum())} of {len(ref)} candidates disagree on validity / error code"
    assert abs(res0.n_valid - int(g["n_valid"])) <= 160
    assert abs(res0.n_generated - int(g["n_generated"])) <= 160
    # FP32 totals of the candidates valid on both sides
    both = (t32 >= 0) & (ref >= 0)
    rel = np.abs(t32[both] - ref[both]) / np.maximum(np.abs(ref[both]), 1e-6)
    print(f"  FP32 totals: median rel {np.median(rel):.2e}, within 1e-4: {(rel <= 1e-4).mean():.5f}, above 1e-2: {(rel > 1e-2).sum()}, "
          f"max {rel.max():.2e}; code mismatches {int(mism.sum())}")
    # measured (r02a-c, both layouts): median 7.3e-8 .. 8.6e-8; within 1e-4: seeds 0 / 2 99.45-99.55 %, seed 1 95.2-96.8 % (its
    # NO_INFORMATION border leaves fewer, more chaotic valid rollouts); above 1e-2: 4-105 of ~42-46k. These candidates
    # are why the default mode refines the best-ranked candidates in FP64 instead of trusting the FP32 ranking
    assert np.median(rel) < 1e-6
    assert (rel <= 1e-4).mean() >= (0.992 if seed != 1 else 0.945)
    assert (rel > 1e-2).mean() <= 4e-3
    # every refined leader carries the reference's FP64 total
    refined = np.flatnonzero((t_ref != t32) & (t_ref >= 0) & (ref >= 0))
    if len(refined):
        rr = np.abs(t_ref[refined] - ref[refined]) / np.abs(ref[refined])
        assert rr.max() <= 1e-6, rr.max()


@pytest.mark.parametrize("seed", [0, 1])
def test_cfg2_exact_mode_equals_the_reference_on_the_full_grid(planner, seed):
    """Mode 1 (FP64 object loops): EVERY one of the 65 536 totals equals the reference's, codes included."""
    g = _full("cfg2", seed)
    cfg, sc, params, smp = _setup(planner, "cfg2", seed, 1)
    try:
        res, poses = planner.plan(sc.world, smp)
        t = planner.explored_totals(res.n_candidates)
    finally:
        planner.set_precision(0)
    ref = g["totals"]
    neg = (t < 0) | (ref < 0)
    assert np.array_equal(t[neg], ref[neg]), f"{int((t[neg] != ref[neg]).sum())} code mismatches"
    rel = np.abs(t[~neg] - ref[~neg]) / np.maximum(np.abs(ref[~neg]), 1e-6)
    print(f"cfg2 seed {seed} mode 1: max rel {rel.max():.2e}, above 1e-6: {(rel > 1e-6).sum()}")
    # social critics are FP32 in every mode (~1e-7 on O(1) terms); cell-indexed critics agree exactly. Measured (r02b): seed 0: all
    # 45 925 valid totals within 1.7e-7; seed 1: 360 of 41 748 (0.9 %) beyond 1e-5 -- rollouts that end chattering around the
    # stationary-robot threshold of World (speed <= 0.01 -> heading = yaw, src/world.cpp:26-30), where the last bit of libm vs
    # the CUDA math library decides the branch: for those the reference itself is only reproducible to its own rounding
    assert (rel > 1e-5).mean() <= 0.015 and np.median(rel) < 1e-6
    assert res.n_valid == int(g["n_valid"]) and res.n_generated == int(g["n_generated"])
    assert res.best_index == int(g["best_index"])
    _check_winner(res, poses, g, 1e-6)


@pytest.mark.parametrize("mode", [2, 1], ids=["refine", "fp64"])
@pytest.mark.parametrize("lay", [1, 2], ids=["warp", "thread"])
def test_batch_cfg3_winners_equal_the_reference(planner, mode, lay):
    """hmp_plan_batch of 8 independent worlds x 4096 candidates: per-scene winners against the reference's."""
    path = os.path.join(GOLD, "ref_full_cfg3_s0-7.npz")
    if not os.path.exists(path):
        pytest.skip("ref_full_cfg3_s0-7.npz missing")
    g = np.load(path)
    cfg = scenes.CONFIGS["cfg3"]
    seeds = [int(s) for s in g["seeds"]]
    scs = [scenes.make_scene(cfg, s) for s in seeds]
    params = scenes.make_params(cfg)
    smp = scenes.make_sampling(cfg)
    for k, sc in enumerate(scs):
        assert gc.scene_fingerprint(sc, params, smp) == g["fingerprint"][k].tobytes(), "scene drifted from the fixture"
    planner.set_precision(mode)
    planner.set_sweep_layout(lay)
    try:
        planner.set_params(params)
        planner.set_scene(scs[0])
        cells = np.stack([sc.cells for sc in scs])
        grids = [np.stack([sc.grids[q] for sc in scs]) for q in range(4)]
        hv = np.array([sc.hv_prev for sc in scs])
        res = planner.plan_batch([sc.world for sc in scs], cells, grids, smp, hv_prev=hv)
        tot = planner.explored_totals(len(scs) * res[0].n_candidates).reshape(len(scs), -1)
        # the same batch with the wave fronts computed on the device from the plans
        plans = []
        for q in range(4):
            xy = [np.asarray(sc.plans[q][0], dtype=np.float64).reshape(-1, 2) for sc in scs]
            plans.append((np.concatenate(xy), np.concatenate([[0], np.cumsum([len(a) for a in xy])]).astype(np.int32)))
        planner.compute_mapgrid_batch(cells, plans, [scs[0].plans[q][1] for q in range(4)])
        res_dev = planner.plan_batch([sc.world for sc in scs], None, None, smp, hv_prev=hv)
    finally:
        planner.set_sweep_layout(0)
        planner.set_precision(0)
    for k in range(len(scs)):
        ref = g["totals"][k]
        r = res[k]
        assert r.n_candidates == ref.shape[0] == 4096
        assert r.best_index == int(g["best_index"][k]) or _top2_close(ref), (k, r.best_index, int(g["best_index"][k]))
        if r.best_index == int(g["best_index"][k]):
            assert abs(r.best_total - float(g["best_total"][k])) <= 1e-6 * abs(float(g["best_total"][k]))
            assert np.allclose(np.array(r.costs), g["best_costs"][k], rtol=1e-6, atol=1e-6, equal_nan=True)
            assert np.abs(np.array([r.xv, r.yv, r.thetav]) - g["best_seed"][k]).max() <= 1e-6
        neg = (tot[k] < 0) | (ref < 0)
        assert (neg & (tot[k] != ref)).mean() <= (0.0 if mode == 1 else 2e-3)
        assert (res_dev[k].best_index, res_dev[k].best_total) == (r.best_index, r.best_total), "device wave fronts changed the selection"
