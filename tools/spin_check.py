"""Developer tool: the crowd-stress worlds (cfg2, seeds 0-9) with a robot that is turning (base velocity (0.5, 0, -1.3) and
(0.2, 0, 2.0)): unreliable-leader count of the mode-2 pass, whether mode 2 WITHOUT escalation returns the exact mode's winner,
and what the default (escalation from 24) returns and costs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, scenes  # noqa: E402

cfg = scenes.CONFIGS["cfg2"]
pl = Planner(0)
pl.set_params(scenes.make_params(cfg))
smp = scenes.make_sampling(cfg)
for vel in ((0.5, 0.0, -1.3), (0.2, 0.0, 2.0)):
    miss_off = miss_on = esc = 0
    for seed in range(10):
        sc = scenes.make_scene(cfg, seed, base_vel=vel)
        pl.set_scene(sc)
        pl.set_precision(1)
        ex, _ = pl.plan(sc.world, smp, want_poses=False)
        t64 = pl.explored_totals(ex.n_candidates)
        v = np.sort(t64[t64 >= 0])
        close = len(v) > 1 and (v[1] - v[0]) <= 1e-4 * abs(v[0])
        pl.set_precision(2)
        pl.set_escalation(0)
        r0, _ = pl.plan(sc.world, smp, want_poses=False)
        unrel = pl.last_unreliable_leaders()
        pl.set_escalation(24)
        t0 = time.perf_counter()
        r1, _ = pl.plan(sc.world, smp, want_poses=False)
        ms = 1e3 * (time.perf_counter() - t0)
        m0 = not (r0.best_index == ex.best_index or close)
        m1 = not (r1.best_index == ex.best_index or close)
        miss_off += m0
        miss_on += m1
        esc += pl.last_escalated()
        print(f"vel {vel} seed {seed}: unreliable {unrel:4d} of {pl.last_num_leaders()} | exact {ex.best_index} mode-2 {r0.best_index} "
              f"{'MISS' if m0 else 'ok'} | default {r1.best_index} {'MISS' if m1 else 'ok'} escalated {pl.last_escalated()} ({ms:.1f} ms) | valid {ex.n_valid}")
    print(f"vel {vel}: misses without escalation {miss_off}, with {miss_on}, escalated {esc} of 10")
