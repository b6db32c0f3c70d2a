#!/usr/bin/env python
"""How wrong can the FP32 ranking be where it matters? For every plan of a 64k-candidate closed-loop replay (and the cfg2
benchmark worlds) the exact FP64 totals (precision mode 1) are compared with the FP32 sweep's totals (mode 0):
rank and relative excess of the TRUE winner in the FP32 ordering, and the FP32 error of the true top-50. This is the
evidence behind the leader-selection rule of mode 2. Runs on the GPU box:

    python tools/selection_hole_stats.py [--plans 120] > gpurun_out/selection_hole.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, scenes, config, replay  # noqa: E402


def stats(t32, t64):
    v64 = np.flatnonzero(t64 >= 0)
    if len(v64) == 0:
        return None
    order = v64[np.argsort(t64[v64], kind="stable")]
    w = int(order[0])
    v32 = t32 >= 0
    best32 = t32[v32].min() if v32.any() else float("nan")
    out = {"winner": w, "winner_total64": float(t64[w]), "winner_total32": float(t32[w]),
           "winner_rank32": int((t32[v32] < t32[w]).sum()) if t32[w] >= 0 else -1,
           "winner_excess32": float((t32[w] - best32) / best32) if t32[w] >= 0 else None,
           "best32_err": float((best32 - t64[np.flatnonzero(v32)[np.argmin(t32[v32])]]) / best32)}
    top = order[:50]
    both = top[(t32[top] >= 0)]
    rel = (t32[both] - t64[both]) / t64[both]
    out["top50_max_abs_rel"] = float(np.abs(rel).max()) if len(both) else None
    out["top50_invalid32"] = int(len(top) - len(both))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plans", type=int, default=120)
    a = ap.parse_args()
    pl = Planner(0)
    rows = []

    def on_plan(params, sc, smp, res):
        tots = {}
        for mode in (1, 0):
            pl.set_precision(mode)
            r, _ = pl.plan(sc.world, smp, want_poses=False)
            tots[mode] = pl.explored_totals(r.n_candidates)
        pl.set_precision(2)
        s = stats(tots[0], tots[1])
        if s is not None:
            s["mode2_pick"] = int(res.best_index)
            s["mode2_ok"] = bool(res.best_index == s["winner"] or abs(res.best_total - s["winner_total64"]) <= 1e-4 * s["winner_total64"])
            s["leaders"] = pl.last_num_leaders()
            rows.append(s)
        return True

    pl.set_precision(2)
    for lay in (1, 2):
        pl.set_sweep_layout(lay)
        replay.run_replay(pl, n_cycles=a.plans + 40, sampling_axes=config.SAMPLING_64K, on_plan=on_plan, on_plan_every=1)
    pl.set_sweep_layout(0)
    # the benchmark worlds
    cfg = scenes.CONFIGS["cfg2"]
    bench_rows = []
    for seed in range(10):
        sc = scenes.make_scene(cfg, seed)
        pl.set_params(scenes.make_params(cfg))
        pl.set_scene(sc)
        smp = scenes.make_sampling(cfg)
        tots = {}
        for mode in (1, 0, 2):
            pl.set_precision(mode)
            r, _ = pl.plan(sc.world, smp, want_poses=False)
            tots[mode] = pl.explored_totals(r.n_candidates)
            if mode == 2:
                res = r
        s = stats(tots[0], tots[1])
        s["seed"] = seed
        s["mode2_ok"] = bool(res.best_index == s["winner"])
        s["leaders"] = pl.last_num_leaders()
        bench_rows.append(s)
    ex = np.array([r["winner_excess32"] for r in rows if r["winner_excess32"] is not None])
    rk = np.array([r["winner_rank32"] for r in rows])
    summary = {"replay_plans": len(rows), "winner_rank32_max": int(rk.max()), "winner_rank32_gt0": int((rk > 0).sum()),
               "winner_excess32_max": float(ex.max()), "winner_excess32_gt_2pct": int((ex > 0.02).sum()),
               "winner_excess32_gt_1pct": int((ex > 0.01).sum()), "mode2_wrong": int(sum(1 for r in rows if not r["mode2_ok"])),
               "top50_max_abs_rel_max": float(max(r["top50_max_abs_rel"] or 0 for r in rows)),
               "bench_winner_rank32": [r["winner_rank32"] for r in bench_rows], "bench_mode2_ok": [r["mode2_ok"] for r in bench_rows]}
    print(json.dumps({"summary": summary, "replay": rows, "bench": bench_rows}))
    pl.close()


if __name__ == "__main__":
    main()
