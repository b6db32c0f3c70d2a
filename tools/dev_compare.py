"""Developer tool: GPU path vs CPU oracle on one synthetic config, with per-term error statistics."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from humap_local_planner_b200 import Planner, scenes  # noqa: E402
from humap_local_planner_b200.capi import COST_NAMES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg0")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--n", type=int, default=72, help="candidates compared against the oracle")
    ap.add_argument("--fis", type=int, default=1)
    ap.add_argument("--forces", type=int, default=-1)
    ap.add_argument("--worst", type=int, default=0)
    ap.add_argument("--precise", type=int, default=0)
    a = ap.parse_args()
    cfg = scenes.CONFIGS[a.cfg]
    sc = scenes.make_scene(cfg, a.seed)
    params = scenes.make_params(cfg, fis=bool(a.fis))
    smp = scenes.make_sampling(cfg)
    pl = Planner(0)
    pl.set_params(params)
    pl.set_precision(bool(a.precise))
    pl.set_scene(sc)
    t = time.time()
    res, poses = pl.plan(sc.world, smp)
    t1 = time.time() - t
    t = time.time()
    res, poses = pl.plan(sc.world, smp)
    t2 = time.time() - t
    C = res.n_candidates
    print(f"{a.cfg} seed {a.seed}: C={C} T={pl.num_steps()} first call {t1*1e3:.1f} ms, second {t2*1e3:.1f} ms, gpu_ms {res.gpu_ms:.3f} "
          f"(select {res.gpu_ms_select:.3f}) best {res.best_index} total {res.best_total:.6f} gen {res.n_generated} valid {res.n_valid}")
    totals = pl.explored_totals(C)
    idx = np.unique(np.linspace(0, C - 1, min(a.n, C)).astype(np.int32))
    ex = pl.explain(idx, with_forces=True)
    T = pl.num_steps()
    o_tot, o_costs, o_poses, o_seeds, o_np = [], [], [], [], []
    t = time.time()
    for i in idx:
        r = ob.plan(params, sc, smp, cand_range=(int(i), int(i) + 1), forces_candidate=int(i) if i == a.forces else -1)
        o_tot.append(r["totals"][i]); o_costs.append(r["costs"][i]); o_poses.append(r["poses"][i]); o_seeds.append(r["seeds"][i]); o_np.append(r["n_poses"][i])
        if i == a.forces:
            k = list(idx).index(i)
            df = ex["forces"][k] - r["forces"]
            print("forces diff per step (max abs over comps):", np.round(np.abs(df).max(axis=1), 6))
            print("gpu forces step0", ex["forces"][k][0], "\norc forces step0", r["forces"][0])
    print(f"oracle: {len(idx)} candidates in {time.time()-t:.2f} s")
    o_tot = np.array(o_tot); o_costs = np.array(o_costs); o_poses = np.array(o_poses); o_np = np.array(o_np)
    g_tot = totals[idx]
    both_neg = (g_tot < 0) & (o_tot < 0)
    code_match = (g_tot == o_tot) | ~both_neg
    print("negative-code agreement:", int((both_neg & (g_tot == o_tot)).sum()), "/", int(both_neg.sum()), "one-sided neg:", int(((g_tot < 0) ^ (o_tot < 0)).sum()))
    gen = (o_np == T) & (ex["n_poses"] == T)
    perr = np.abs(ex["poses"][gen] - o_poses[gen])
    perr[..., 2] = np.abs((perr[..., 2] + np.pi) % (2 * np.pi) - np.pi)
    if gen.any():
        print(f"pose err: max xy {perr[..., :2].max():.3e} max yaw {perr[..., 2].max():.3e}; frac cands > 1e-4: {(perr.max(axis=(1,2)) > 1e-4).mean():.4f}")
    if gen.any() and a.worst:
        gi = np.where(gen)[0]
        w = gi[np.argmax(perr.max(axis=(1, 2)))]
        ci = int(idx[w])
        r = ob.plan(params, sc, smp, cand_range=(ci, ci + 1), forces_candidate=ci)
        pe = np.abs(ex["poses"][w] - r["poses"][ci])
        df = ex["forces"][w] - r["forces"]
        print(f"worst candidate {ci}: per-step [pose err xy, yaw | dF int, dyn, stat, human | F gpu dyn, stat, human]")
        for i in range(T):
            print(f"  {i:2d} {pe[i,:2].max():.2e} {pe[i,2]:.2e} | {np.abs(df[i,0:2]).max():.2e} {np.abs(df[i,2:4]).max():.2e} {np.abs(df[i,4:6]).max():.2e} {np.abs(df[i,6:8]).max():.2e} | "
                  f"{np.hypot(*ex['forces'][w][i,2:4]):.3f} {np.hypot(*ex['forces'][w][i,4:6]):.3f} {np.hypot(*ex['forces'][w][i,6:8]):.3f}")
    for k, name in enumerate(COST_NAMES):
        g, o = ex["costs"][:, k], o_costs[:, k]
        m = np.isfinite(g) & np.isfinite(o)
        if not m.any():
            print(f"  {name:16s}: not evaluated"); continue
        rel = np.abs(g[m] - o[m]) / np.maximum(np.abs(o[m]), 1e-9)
        absd = np.abs(g[m] - o[m])
        bad = (rel > 1e-4) & (absd > 1e-6)
        print(f"  {name:16s}: n={int(m.sum()):5d} max rel {rel.max():.3e} max abs {absd.max():.3e} frac>1e-4 {bad.mean():.4f} nan-mismatch {int((np.isfinite(g) ^ np.isfinite(o)).sum())}")
    v = (g_tot >= 0) & (o_tot >= 0)
    if v.any():
        rel = np.abs(g_tot[v] - o_tot[v]) / np.maximum(np.abs(o_tot[v]), 1e-9)
        print(f"totals: n={int(v.sum())} max rel {rel.max():.3e} frac>1e-4 {(rel > 1e-4).mean():.4f}")
    pl.close()


if __name__ == "__main__":
    main()
