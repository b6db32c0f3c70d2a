"""Developer tool: FP32-sweep accuracy of the CUDA path against the oracle -- pose error percentiles and totals per config.
Select the library with HMP_LIB=... to compare two builds on the same box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from humap_local_planner_b200 import Planner, scenes  # noqa: E402


def main():
    pl = Planner(0)
    cases = [("cfg0", s, 72) for s in range(8)] + [("cfg1", 0, 256), ("cfg1", 1, 256), ("cfg2", 0, 128), ("cfg2", 1, 128), ("cfg2", 2, 128)]
    for name, seed, n in cases:
        cfg = scenes.CONFIGS[name]
        sc = scenes.make_scene(cfg, seed)
        params = scenes.make_params(cfg)
        smp = scenes.make_sampling(cfg)
        pl.set_precision(0)
        pl.set_params(params)
        pl.set_scene(sc)
        res, _ = pl.plan(sc.world, smp)
        C = res.n_candidates
        T = pl.num_steps()
        idx = np.unique(np.linspace(0, C - 1, min(n, C)).astype(np.int32))
        ex = pl.explain(idx)
        tot = pl.explored_totals(C)[idx]
        o = ob.plan_sampled(params, sc, smp, idx)
        both = (ex["n_poses"] == T) & (o["n_poses"] == T)
        e = np.abs(ex["poses"][both][..., :2] - o["poses"][both][..., :2]).max(axis=(1, 2))
        v = (tot >= 0) & (o["totals"] >= 0)
        rel = np.abs(tot[v] - o["totals"][v]) / np.abs(o["totals"][v])
        print(f"{name} s{seed}: poses med {np.median(e):.2e} p90 {np.percentile(e, 90):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e} "
              f"within1e-4 {(e <= 1e-4).mean():.3f} | totals rel med {np.median(rel):.1e} p90 {np.percentile(rel, 90):.1e} >1e-3: {(rel > 1e-3).mean():.3f} "
              f"| best {res.best_index} {res.best_total:.6f} ms {res.gpu_ms_select:.3f}")
    pl.close()


if __name__ == "__main__":
    main()
