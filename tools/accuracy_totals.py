"""FP32 sweep totals of both work layouts against the oracle: cfg1 over the full 16k grid, cfg2 on an evenly spaced sample.
Prints, per layout, the agreement of the invalid codes and the distribution of the relative error of the valid totals.
usage (GPU box): python tools/accuracy_totals.py [n_cfg2_sample]"""
import concurrent.futures as cf
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from humap_local_planner_b200 import Planner, scenes  # noqa: E402

n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
pl = Planner(0)
pl.set_precision(0)
out = {}
for name, seed in (("cfg1", 0), ("cfg2", 0)):
    cfg = scenes.CONFIGS[name]
    sc = scenes.make_scene(cfg, seed)
    params = scenes.make_params(cfg)
    smp = scenes.make_sampling(cfg)
    pl.set_params(params)
    pl.set_scene(sc)
    tot = {}
    for lay in (1, 2):
        pl.set_sweep_layout(lay)
        r, _ = pl.plan(sc.world, smp, want_poses=False)
        tot[lay] = pl.explored_totals(r.n_candidates)
    C = len(tot[1])
    if name == "cfg1":
        idx = np.arange(C)
        ref = ob.plan_all_threaded(params, sc, smp)
    else:
        idx = np.unique(np.linspace(0, C - 1, n2).astype(np.int64))
        chunks = np.array_split(idx, 4 * (os.cpu_count() or 1))
        with cf.ThreadPoolExecutor(os.cpu_count() or 1) as ex:
            parts = list(ex.map(lambda ch: ob.plan_sampled(params, sc, smp, ch)["totals"] if len(ch) else np.zeros(0), chunks))
        ref = np.concatenate(parts)
    for lay in (1, 2):
        g = tot[lay][idx]
        v = (g >= 0) & (ref >= 0)
        rel = np.abs(g[v] - ref[v]) / np.maximum(np.abs(ref[v]), 1e-6)
        out[f"{name}_layout{lay}"] = {
            "candidates": int(len(idx)), "both_valid": int(v.sum()), "validity_agrees": float(((g >= 0) == (ref >= 0)).mean()),
            "codes_equal_where_both_invalid": float((g[(g < 0) & (ref < 0)] == ref[(g < 0) & (ref < 0)]).mean()) if ((g < 0) & (ref < 0)).any() else None,
            "rel_err_median": float(np.median(rel)), "rel_err_p90": float(np.percentile(rel, 90)), "rel_err_p99": float(np.percentile(rel, 99)),
            "share_within_1e-4": float((rel <= 1e-4).mean()), "share_within_1e-3": float((rel <= 1e-3).mean()),
            "share_above_1e-2": float((rel > 1e-2).mean())}
print(json.dumps(out, indent=1))
