"""Developer tool: 64k-candidate closed-loop replay; for the first plans whose TRUE winner (FP64 sweep) carries an FP32 total more
than 2 % above the FP32 best, print where its FP32 rollout leaves the FP64 one: per-step pose error, robot speed, forces,
and the critics that differ."""
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


def on_plan(params, sc, smp, res):
    if shown[0] >= LIMIT:
        return True
    tots, exs = {}, {}
    for mode in (1, 0):
        pl.set_precision(mode)
        r, _ = pl.plan(sc.world, smp, want_poses=False)
        tots[mode] = pl.explored_totals(r.n_candidates)
    v64 = np.flatnonzero(tots[1] >= 0)
    if len(v64) == 0:
        pl.set_precision(2)
        return True
    w = int(v64[np.argmin(tots[1][v64])])
    b32 = tots[0][tots[0] >= 0].min()
    if tots[0][w] < 0 or (tots[0][w] - b32) / b32 <= 0.02:
        pl.set_precision(2)
        return True
    shown[0] += 1
    for mode in (1, 0):
        pl.set_precision(mode)
        pl.plan(sc.world, smp, want_poses=False)
        exs[mode] = pl.explain([w], with_forces=True)
    pl.set_precision(2)
    T = pl.num_steps()
    dt = params.general.sim_time / T
    p64, p32 = exs[1]["poses"][0], exs[0]["poses"][0]
    print(f"=== winner {w}: total64 {tots[1][w]:.4f} total32 {tots[0][w]:.4f} (FP32 best {b32:.4f}); robot vel ({sc.world.vel_x:.3f}, {sc.world.vel_th:.3f})")
    d = exs[0]["costs"][0] - exs[1]["costs"][0]
    print("   critics (fp32 - fp64):", {COST_NAMES[k]: (round(float(exs[0]['costs'][0][k]), 4), round(float(exs[1]['costs'][0][k]), 4)) for k in range(14) if abs(d[k]) > 1e-4})
    print("   step | pose err xy / yaw | speed64 speed32 | F64 (int dyn stat hum) | F32")
    for i in range(T):
        e = np.abs(p32[i, :2] - p64[i, :2]).max()
        ey = abs((p32[i, 2] - p64[i, 2] + np.pi) % (2 * np.pi) - np.pi)
        s64 = np.hypot(*(p64[min(i + 1, T - 1), :2] - p64[max(i, 0) if i + 1 < T else i - 1, :2])) / dt
        s32 = np.hypot(*(p32[min(i + 1, T - 1), :2] - p32[max(i, 0) if i + 1 < T else i - 1, :2])) / dt
        f64, f32 = exs[1]["forces"][0][i], exs[0]["forces"][0][i]
        print(f"   {i:3d} {e:9.2e} {ey:9.2e} | {s64:8.5f} {s32:8.5f} | " + " ".join(f"{v:9.3f}" for v in f64) + " | " + " ".join(f"{v:9.3f}" for v in f32))
    return True


pl.set_precision(2)
replay.run_replay(pl, n_cycles=200, sampling_axes=config.SAMPLING_64K, on_plan=on_plan, on_plan_every=1)
