#!/usr/bin/env python
"""Times one synthetic world of a configuration in a given precision mode (resident inputs, CUDA events) and prints the
counters of the plan -- the unit of work behind a per-seed line of bench.py, small enough to put under ncu:

    python tools/time_seed.py --cfg cfg2 --seed 2 --precise 1 --reps 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--precise", type=int, default=2)
ap.add_argument("--layout", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = scenes.CONFIGS[a.cfg]
sc = scenes.make_scene(cfg, seed=a.seed)
pl = Planner(0)
pl.set_precision(a.precise)
pl.set_sweep_layout(a.layout)
pl.set_params(scenes.make_params(cfg))
pl.set_scene(sc)
res, _ = pl.plan(sc.world, scenes.make_sampling(cfg))
ms = []
for _ in range(a.reps):
    r = pl.replan_resident()[0]
    ms.append((round(r.gpu_ms, 3), round(r.gpu_ms_select, 3)))
print(json.dumps({"cfg": a.cfg, "seed": a.seed, "precise": a.precise, "cycle_sweep_ms": ms, "best_index": int(res.best_index),
                  "best_total": float(res.best_total), "n_generated": int(res.n_generated), "n_valid": int(res.n_valid),
                  "leaders": pl.last_num_leaders(), "sweep_mode": pl.last_sweep_mode()}))
pl.close()
