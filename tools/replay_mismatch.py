"""Developer tool: closed-loop replay in FP32 mode; for every checked cycle whose selection differs from the oracle's, print
the per-critic values (device vs oracle) of both candidates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from humap_local_planner_b200 import Planner, replay  # noqa: E402
from humap_local_planner_b200.capi import COST_NAMES  # noqa: E402

pl = Planner(0)
pl.set_precision(0)
stats = {"checked": 0, "mismatch": 0}


def check(params, sc, smp, res):
    stats["checked"] += 1
    ref = ob.plan(params, sc, smp, early_exit=False)
    tot = ref["totals"]
    rb = ref["result"].best_index
    if rb == res.best_index:
        return True
    stats["mismatch"] += 1
    ex = pl.explain([rb, res.best_index])
    g = pl.explored_totals(res.n_candidates)
    print(f"--- oracle best {rb} ({tot[rb]:.5f}; device total {g[rb]:.5f}) device best {res.best_index} "
          f"(oracle total {tot[res.best_index]:.5f}; device {g[res.best_index]:.5f})")
    for j, c in enumerate((rb, res.best_index)):
        d = ex["costs"][j] - ref["costs"][c]
        bad = [(COST_NAMES[k], float(ex["costs"][j][k]), float(ref["costs"][c][k])) for k in range(14) if abs(d[k]) > 1e-4 * max(1, abs(ref["costs"][c][k]))]
        pe = np.abs(ex["poses"][j] - ref["poses"][c]).max()
        print(f"   cand {c}: max pose err {pe:.2e}; critics that differ: {bad}")
    return False


log = replay.run_replay(pl, n_cycles=900, on_plan=check, on_plan_every=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
print(stats)
