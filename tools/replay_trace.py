"""Developer tool: closed-loop replay (FP32 sweep); at the first checked cycle where some candidate's device poses are more
than `thr` metres off the oracle's, print the per-step force components of that candidate on both sides."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from humap_local_planner_b200 import Planner, replay  # noqa: E402

pl = Planner(0)
pl.set_precision(0)
thr = float(sys.argv[1]) if len(sys.argv) > 1 else 5e-3
found = []


def check(params, sc, smp, res):
    if found:
        return True
    ref = ob.plan(params, sc, smp, early_exit=False)
    C = res.n_candidates
    ex = pl.explain(np.arange(C, dtype=np.int32), with_forces=True)
    T = pl.num_steps()
    both = (ex["n_poses"] == T) & (ref["n_poses"] == T)
    e = np.where(both, np.abs(ex["poses"][..., :2] - ref["poses"][..., :2]).max(axis=(1, 2)), 0.0)
    c = int(np.argmax(e))
    if e[c] < thr:
        return True
    found.append(c)
    r = ob.plan(params, sc, smp, cand_range=(c, c + 1), forces_candidate=c)
    gf, of = ex["forces"][c], r["forces"]
    print(f"candidate {c}: max pose err {e[c]:.3e}; n_people {sc.world.n_people} n_obst {sc.world.n_obstacles}")
    print("step | pose err | int dyn stat human (device)            | (oracle)")
    for i in range(T):
        pe = np.abs(ex["poses"][c][i, :2] - ref["poses"][c][i, :2]).max()
        print(f"{i:3d} {pe:9.2e} | " + " ".join(f"{v:10.4f}" for v in gf[i]) + " | " + " ".join(f"{v:10.4f}" for v in of[i]))
    return True


replay.run_replay(pl, n_cycles=900, on_plan=check, on_plan_every=5)
