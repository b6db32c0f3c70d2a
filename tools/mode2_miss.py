"""Developer tool: 64k-candidate closed-loop replay (1000 cycles); for every plan whose mode-2 winner differs from the exact
mode's, print the rank of the exact winner in the FP32 ordering, the FP32 / FP64 totals of both candidates and the critics that
differ between the FP32 sweep and the FP64 evaluation of the exact winner."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humap_local_planner_b200 import Planner, replay, config  # noqa: E402
from humap_local_planner_b200.capi import COST_NAMES  # noqa: E402

pl = Planner(0)
lay = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = [0]


def on_plan(params, sc, smp, res):
    n[0] += 1
    pl.set_precision(1)
    exact, _ = pl.plan(sc.world, smp, want_poses=False)
    t64 = pl.explored_totals(exact.n_candidates)
    pl.set_precision(2)
    if exact.best_index == res.best_index:
        return True
    v = np.sort(t64[t64 >= 0])
    close = len(v) > 1 and (v[1] - v[0]) <= 1e-4 * abs(v[0])
    pl.set_precision(0)
    r0, _ = pl.plan(sc.world, smp, want_poses=False)
    t32 = pl.explored_totals(r0.n_candidates)
    pl.set_precision(2)
    w, m = exact.best_index, res.best_index
    valid32 = t32 >= 0
    rank = int((t32[valid32] < t32[w]).sum()) if t32[w] >= 0 else -1
    print(f"plan {n[0] - 1}: exact winner {w} (t64 {t64[w]:.6f}, t32 {t32[w]:.6f}, FP32 rank {rank}) | mode-2 winner {m} (t64 {t64[m]:.6f}, t32 {t32[m]:.6f}, "
          f"reported {res.best_total:.6f}) | FP32 best {t32[valid32].min():.6f} | top-2 close {close} | leaders {pl.last_num_leaders()} "
          f"valid32 {int(valid32.sum())} valid64 {int((t64 >= 0).sum())} | robot vel ({sc.world.vel_x:.3f}, {sc.world.vel_th:.3f})")
    ex = {}
    for mode in (1, 0):
        pl.set_precision(mode)
        pl.plan(sc.world, smp, want_poses=False)
        ex[mode] = pl.explain([w, m])
    pl.set_precision(2)
    for j, c in enumerate((w, m)):
        d = ex[0]["costs"][j] - ex[1]["costs"][j]
        print(f"   candidate {c}: critics (explain fp32, fp64) that differ:",
              {COST_NAMES[k]: (round(float(ex[0]['costs'][j][k]), 5), round(float(ex[1]['costs'][j][k]), 5)) for k in range(14) if abs(d[k]) > 1e-6 or np.isnan(d[k])})
    if os.environ.get("TRACE") and n[0] - 1 == int(os.environ["TRACE"]):
        exs = {}
        for mode in (1, 0):
            pl.set_precision(mode)
            pl.plan(sc.world, smp, want_poses=False)
            exs[mode] = pl.explain([w], with_forces=True)
        pl.set_precision(2)
        p64, p32 = exs[1]["poses"][0], exs[0]["poses"][0]
        T = p64.shape[0]
        dt = params.general.sim_time / T
        print("   step | pose err xy / yaw | speed64 speed32 | yawrate64 yawrate32 | F64 (int dyn stat hum) x,y | F32")
        for i in range(T - 1):
            e = np.abs(p32[i, :2] - p64[i, :2]).max()
            ey = abs((p32[i, 2] - p64[i, 2] + np.pi) % (2 * np.pi) - np.pi)
            s64 = np.hypot(*(p64[i + 1, :2] - p64[i, :2])) / dt
            s32 = np.hypot(*(p32[i + 1, :2] - p32[i, :2])) / dt
            w64 = ((p64[i + 1, 2] - p64[i, 2] + np.pi) % (2 * np.pi) - np.pi) / dt
            w32 = ((p32[i + 1, 2] - p32[i, 2] + np.pi) % (2 * np.pi) - np.pi) / dt
            f64, f32 = exs[1]["forces"][0][i], exs[0]["forces"][0][i]
            print(f"   {i:3d} {e:9.2e} {ey:9.2e} | {s64:8.5f} {s32:8.5f} | {w64:8.4f} {w32:8.4f} | " + " ".join(f"{v:8.2f}" for v in f64) + " | " + " ".join(f"{v:8.2f}" for v in f32))
    return True


pl.set_precision(2)
pl.set_sweep_layout(lay)
replay.run_replay(pl, n_cycles=1000, sampling_axes=config.SAMPLING_64K, on_plan=on_plan, on_plan_every=1)
print("plans", n[0])
