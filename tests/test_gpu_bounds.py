"""The small parity cases once more through the bounds-checked debug build (lib/libhmp_planner_check.so, -DHMP_BOUNDS_CHECK):
every index into shared / global memory that is computed from scene data is asserted on the device (HMP_CHECK in
csrc/hmp_kernels.cu / hmp_sweep_tpc.inl). compute-sanitizer is not available on the GPU pool; this is the substitute.
Runs in a subprocess so that the regular library of this process is not replaced."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import json, sys
import numpy as np
sys.path.insert(0, %(root)r)
from humap_local_planner_b200 import Planner, scenes, config
from humap_local_planner_b200.capi import HmpEquisampled
out = {}
pl = Planner(0)
for name, seed, lay, mode in (("cfg0", 0, 1, 2), ("cfg0", 1, 2, 2), ("cfg0", 3, 1, 1), ("cfg1", 0, 2, 2), ("cfg1", 1, 1, 0), ("cfg2", 1, 2, 2),
                             ("cfg1", 2, 2, 1), ("cfg2", 0, 2, 1)):   # ... and the FP64 thread-per-candidate sweep (r02zz)
    cfg = scenes.CONFIGS[name]
    sc = scenes.make_scene(cfg, seed)
    pl.set_precision(mode); pl.set_sweep_layout(lay)
    pl.set_params(scenes.make_params(cfg)); pl.set_scene(sc)
    smp = scenes.make_sampling(cfg)
    if name == "cfg2":   # a 4k subgrid of the crowd-stress world keeps the debug build quick
        smp = config.make_sampling(config.SAMPLING_4K)
    res, poses = pl.plan(sc.world, smp)
    ex = pl.explain([0, max(0, res.best_index), res.n_candidates - 1])
    out["%%s-%%d-%%d-%%d" %% (name, seed, lay, mode)] = [res.best_index, res.best_total, res.n_valid, float(np.nansum(ex["costs"]))]
# the second generator of the pool: small social pool + large equisampled pool (the d_block_best overflow of ADVICE r1)
cfg = scenes.CONFIGS["cfg0"]; sc = scenes.make_scene(cfg, 0)
pl.set_precision(2); pl.set_sweep_layout(0); pl.set_params(scenes.make_params(cfg)); pl.set_scene(sc)
eq = HmpEquisampled(); eq.enabled, eq.vx_samples, eq.vy_samples, eq.vth_samples, eq.min_vel_x, eq.continued_acceleration = 1, 5, 1, 40, 0.1, 1
pl.set_equisampled(eq)
res, _ = pl.plan(sc.world, scenes.make_sampling(cfg))
out["equi"] = [res.best_index, res.best_total, res.n_candidates, res.n_social]
pl.set_equisampled(None)
# device wave fronts, cost cloud, a batch of four worlds
for g, (plan, lg) in enumerate(sc.plans):
    pl.compute_mapgrid(g, plan, lg, sc.hv_prev[g])
res, _ = pl.plan(sc.world, scenes.make_sampling(cfg))
cloud, valid = pl.cost_cloud()
out["device-grids"] = [res.best_index, res.best_total, int(valid.sum())]
cfg3 = scenes.CONFIGS["cfg3"]; scs = [scenes.make_scene(cfg3, s) for s in range(4)]
pl.set_params(scenes.make_params(cfg3)); pl.set_scene(scs[0])
rb = pl.plan_batch([s.world for s in scs], np.stack([s.cells for s in scs]), [np.stack([s.grids[q] for s in scs]) for q in range(4)],
                   scenes.make_sampling(cfg3), hv_prev=np.array([s.hv_prev for s in scs]))
out["batch"] = [[r.best_index, r.best_total] for r in rb]
# the batch once more with its 4 x 4 wave fronts computed on the device (shared-memory wave-front kernel, batched form)
plans = []
for q in range(4):
    xy = [np.asarray(s.plans[q][0], dtype=np.float64).reshape(-1, 2) for s in scs]
    plans.append((np.concatenate(xy), np.concatenate([[0], np.cumsum([len(a) for a in xy])]).astype(np.int32)))
pl.compute_mapgrid_batch(np.stack([s.cells for s in scs]), plans, [s_[1] for s_ in scs[0].plans])
rb2 = pl.plan_batch([s.world for s in scs], None, None, scenes.make_sampling(cfg3), hv_prev=np.array([s.hv_prev for s in scs]))
out["batch-device-grids"] = [[r.best_index, r.best_total] for r in rb2]
# the same batch with the second generator of the pool (per-world velocity windows, padded sample lists)
eq.vth_samples, eq.continued_acceleration = 10, 0
pl.set_equisampled(eq)
rb = pl.plan_batch([s.world for s in scs], None, None, scenes.make_sampling(cfg3), hv_prev=np.array([s.hv_prev for s in scs]))
out["batch-equi"] = [[r.best_index, r.best_total, r.n_candidates] for r in rb]
pl.set_equisampled(None)
pl.close()
print("RESULT " + json.dumps(out))
'''


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["HMP_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", SNIPPET % {"root": ROOT}], capture_output=True, text=True, timeout=900, env=env)
    return r


@pytest.mark.gpu
def test_bounds_checked_build_runs_the_small_cases_cleanly():
    from humap_local_planner_b200 import build as b
    lib = b.LIB_CHECK
    if not os.path.exists(lib):
        lib = b.build_library(check=True)
    r = _run(lib)
    assert "HMP_CHECK failed" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    checked = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    r0 = _run(None)
    assert r0.returncode == 0, (r0.stdout + r0.stderr)[-2000:]
    plain = json.loads([ln for ln in r0.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    assert checked == plain     # the assertions change nothing: bit-identical results
    assert checked["batch-device-grids"] == checked["batch"]   # wave fronts on the device == host-computed grids


def test_bounds_checked_build_compiles():
    from humap_local_planner_b200 import build as b
    assert os.path.exists(b.build_library(check=True))
