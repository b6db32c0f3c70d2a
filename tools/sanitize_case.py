"""Small end-to-end run for compute-sanitizer: cfg0 cycle (FP32 + FP64), explain, device wave fronts, debug hooks, batch of 3."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, scenes
cfg = scenes.CONFIGS["cfg0"]
sc = scenes.make_scene(cfg, 1)
params = scenes.make_params(cfg); smp = scenes.make_sampling(cfg)
pl = Planner(0)
for precise in (False, True):
    pl.set_precision(precise)
    pl.set_params(params); pl.set_scene(sc)
    for g, (plan, lg) in enumerate(sc.plans):
        pl.compute_mapgrid(g, plan, lg, 0.0)
    res, poses = pl.plan(sc.world, smp)
    ex = pl.explain(np.arange(0, 72, 5), with_forces=True)
    print(precise, res.best_index, res.best_total, res.n_valid, ex["n_poses"][:4])
pl.set_precision(False)
print(pl.debug_fis(np.random.default_rng(0).uniform(-3, 3, (64, 4)))[:2])
print(pl.debug_footprint_cost(np.array([[0, 0, 0], [4.9, 4.9, 1.0], [1.0, 2.0, -2.0]])))
print(pl.debug_world_to_map(np.array([0.0, -6.0, 4.99]), np.array([0.0, 1.0, -4.99])))
c3 = scenes.CONFIGS["cfg3"]
scs = [scenes.make_scene(c3, 400 + i) for i in range(3)]
p3 = scenes.make_params(c3)
pl.set_params(p3); pl.set_scene(scs[0])
s3 = scenes.config.make_sampling({"an": (-0.5, 1.0, 0.5), "aw": (0.5, 2.0, 0.5)})
r = pl.plan_batch([s.world for s in scs], np.stack([s.cells for s in scs]), [np.stack([s.grids[g] for s in scs]) for g in range(4)], s3)
print([x.best_index for x in r])
pl.close()
print("SANITIZE_CASE_DONE")
