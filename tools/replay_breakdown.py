#!/usr/bin/env python3
"""Where the stock-sampling (72 candidates) replay cycle spends its time: wall clock per C-ABI call of a MOVE cycle
(set_costmap, 4 x compute_mapgrid, set_footprint, plan) and the device times hmp_plan reports, for precision modes 2 and 1."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, replay  # noqa: E402


class Timed:
    def __init__(self, pl):
        self.pl = pl
        self.t = {}

    def __getattr__(self, name):
        f = getattr(self.pl, name)
        if not callable(f):
            return f

        def w(*a, **k):
            t0 = time.perf_counter()
            r = f(*a, **k)
            self.t.setdefault(name, []).append(1e3 * (time.perf_counter() - t0))
            return r
        return w


out = {}
for mode in ((2,) if os.environ.get("CANDIDATES") == "64k" else (2, 1)):
    pl = Planner(0)
    pl.set_precision(mode)
    replay.run_replay(pl, n_cycles=60)
    tp = Timed(pl)
    from humap_local_planner_b200 import config
    axes = config.SAMPLING_64K if os.environ.get("CANDIDATES") == "64k" else None   # CANDIDATES=64k: the 64k-candidate grid
    sel = []
    orig_plan = pl.plan

    def plan_and_note(*a, **k):
        r = orig_plan(*a, **k)
        sel.append((r[0].gpu_ms_select, r[0].gpu_ms))
        return r
    pl.plan = plan_and_note
    log = replay.run_replay(tp, n_cycles=int(os.environ.get("CYCLES", "600")), sampling_axes=axes)
    s = replay.summarize(log)
    n_move = s["move_cycles"]
    row = {"p50_cycle_ms": s["p50_cycle_ms"], "p99_cycle_ms": s["p99_cycle_ms"], "p50_gpu_ms": s["p50_gpu_ms"], "move_cycles": n_move}
    for k, v in tp.t.items():
        v = np.array(v)
        row[k] = {"calls_per_cycle": round(len(v) / max(1, n_move), 2), "p50_ms": float(np.percentile(v, 50)), "sum_per_cycle_ms": float(v.sum() / max(1, n_move))}
    if sel:
        row["p50_sweep_ms"] = float(np.percentile([x[0] for x in sel], 50))
        row["p50_after_sweep_ms"] = float(np.percentile([x[1] - x[0] for x in sel], 50))
        row["leaders_last"] = pl.last_num_leaders()
    out[f"precision{mode}"] = row
    pl.close()
print(json.dumps(out, indent=1))
