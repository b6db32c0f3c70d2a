"""Developer tool: where does the thread-per-candidate sweep differ from the warp-per-candidate one? 64k-candidate closed-loop
replay, FP32 sweeps in both layouts on every plan; for the candidate with the largest difference of the totals: raw critics
of the sweep itself (hmp_debug_sweep_candidate) next to the warp-per-candidate FP32 and FP64 detail passes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, replay, config  # noqa: E402
from humap_local_planner_b200.capi import COST_NAMES  # noqa: E402

pl = Planner(0)
shown = [0]
LIMIT = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hist = []


def on_plan(params, sc, smp, res):
    t = {}
    pl.set_precision(0)
    for lay in (1, 2):
        pl.set_sweep_layout(lay)
        r, _ = pl.plan(sc.world, smp, want_poses=False)
        t[lay] = pl.explored_totals(r.n_candidates)
    both = (t[1] >= 0) & (t[2] >= 0)
    rel = np.where(both, np.abs(t[1] - t[2]) / np.maximum(t[1], 1e-9), 0.0)
    hist.append((float(np.median(rel[both])), float((rel > 1e-3).mean()), float(rel.max()), int(((t[1] < 0) != (t[2] < 0)).sum())))
    # among the 200 best of the warp layout
    top = np.flatnonzero(both)[np.argsort(t[1][both], kind="stable")[:200]]
    c = int(top[np.argmax(rel[top])])
    if rel[c] > 5e-3 and shown[0] < LIMIT:
        shown[0] += 1
        dbg = pl.debug_sweep_candidate(c)     # layout 2 is the one of the last plan
        pl.set_sweep_layout(1)
        pl.plan(sc.world, smp, want_poses=False)
        e32 = pl.explain([c])
        pl.set_precision(1)
        pl.plan(sc.world, smp, want_poses=False)
        e64 = pl.explain([c])
        T = pl.num_steps()
        print(f"=== plan {len(hist)} cand {c}: warp total {t[1][c]:.5f} thread total {t[2][c]:.5f}; robot pose ({sc.world.robot_x:.3f}, {sc.world.robot_y:.3f}, "
              f"{sc.world.robot_yaw:.3f}) vel ({sc.world.vel_x:.3f}, {sc.world.vel_th:.3f})")
        for k in range(14):
            a, b, d = dbg[k], e32["costs"][0][k], e64["costs"][0][k]
            flag = " <--" if abs(a - b) > 1e-4 * max(1.0, abs(b)) else ""
            print(f"    {COST_NAMES[k]:18s} thread-sweep {a:12.6f}  warp-f32 {b:12.6f}  f64 {d:12.6f}{flag}")
        print(f"    seed (x, w): thread {dbg[14]:.6f} {dbg[15]:.6f}  warp-f32 {e32['seeds'][0][0]:.6f} {e32['seeds'][0][2]:.6f}")
        print(f"    last recorded pose warp-f32 {e32['poses'][0][T - 1]} f64 {e64['poses'][0][T - 1]}; thread pose after the last step {dbg[16:19]}")
    pl.set_precision(2)
    pl.set_sweep_layout(0)
    return True


pl.set_precision(2)
replay.run_replay(pl, n_cycles=160, sampling_axes=config.SAMPLING_64K, on_plan=on_plan, on_plan_every=1)
h = np.array(hist)
print("plans", len(h), "median rel diff (median over plans)", np.median(h[:, 0]), "share > 1e-3 (max over plans)", h[:, 1].max(), "max rel", h[:, 2].max(),
      "validity mismatches max", h[:, 3].max())
