"""Developer tool: 64k-candidate closed-loop replay (1000 cycles); per plan the number of UNRELIABLE leaders (FP32 total off by
> 1 % against the FP64 refinement, hmp_last_unreliable_leaders) and whether mode 2 selected the exact mode's winner. Prints, for a
range of escalation thresholds, how many plans would be redone in FP64 and how many misses would remain."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, replay, config  # noqa: E402

pl = Planner(0)
lay = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows = []


def on_plan(params, sc, smp, res):
    unrel, lead = pl.last_unreliable_leaders(), pl.last_num_leaders()
    pl.set_precision(1)
    exact, _ = pl.plan(sc.world, smp, want_poses=False)
    t64 = pl.explored_totals(exact.n_candidates)
    pl.set_precision(2)
    v = np.sort(t64[t64 >= 0])
    close = len(v) > 1 and (v[1] - v[0]) <= 1e-4 * abs(v[0])
    miss = not (res.best_index == exact.best_index or close)
    rows.append((len(rows), int(miss), unrel, lead, abs(sc.world.vel_th)))
    return True


pl.set_precision(2)
pl.set_sweep_layout(lay)
replay.run_replay(pl, n_cycles=1000, sampling_axes=config.SAMPLING_64K, on_plan=on_plan, on_plan_every=1)
a = np.array(rows, dtype=float)
print("plans", len(a), "misses", int(a[:, 1].sum()), "at", [(int(r[0]), int(r[2])) for r in a if r[1]])
print("unreliable leaders per plan: percentiles 50 / 90 / 95 / 99 / max:", [float(np.percentile(a[:, 2], q)) for q in (50, 90, 95, 99, 100)])
for thr in (2, 4, 8, 12, 16, 24, 32, 48, 64):
    esc = a[:, 2] >= thr
    print(f"threshold {thr:3d}: {int(esc.sum()):4d} plans redone ({100.0 * esc.mean():.1f} %), misses left {int((a[:, 1] * ~esc).sum())}")
print(json.dumps({"unreliable_hist": np.bincount(np.minimum(a[:, 2].astype(int), 100)).tolist()}))
