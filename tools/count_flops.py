#!/usr/bin/env python3
"""Instrumented floating-point operation count of the path (SURVEY.md 8d / BASELINE.md 3 asked for it; the W of those documents
is an estimate read off the reference's statements).

The CPU oracle is compiled a second time with `double` replaced by a counting scalar (oracle/counted_double.h,
oracle/hmp_oracle_count.cpp -- the oracle's own source text, bit-identical totals) and run on an evenly spaced sample of the
sampling grid of each benchmark configuration. Output: profiles/r02_flop_count.json, which bench.py reads for
`roofline.algorithmic` (no oracle code runs in the bench for this).

flop = add + mul + div + sqrt + exp/pow + sin/cos/acos + atan2, one each (comparisons and roundings are listed, not summed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_binding as ob  # noqa: E402
from humap_local_planner_b200 import scenes  # noqa: E402

FLOP_OPS = ("add", "mul", "div", "sqrt", "exp", "trig", "atan2")


def survey_w(cfg, fis: bool) -> int:
    """W = 294 + 86 S + 103 D + 185 P + 45 G + (K + 1)(16 V + 6 c) [+ 2900 D with FIS]; K = 8, V = 16, c = 45."""
    S, P, G = cfg.n_obstacles, cfg.n_people, cfg.n_groups
    D = P + G
    return 294 + 86 * S + 103 * D + 185 * P + 45 * G + 9 * (16 * 16 + 6 * 45) + (2900 * D if fis else 0)


def count(cfg_name: str, seed: int, n_sample: int) -> dict:
    cfg = scenes.CONFIGS[cfg_name]
    scene = scenes.make_scene(cfg, seed)
    params = scenes.make_params(cfg)
    sampling = scenes.make_sampling(cfg)
    C = ob.num_candidates(sampling)
    T = ob.num_steps(params, scene.world)
    idx = np.unique(np.linspace(0, C - 1, min(C, n_sample)).astype(np.int64))
    t0 = time.time()
    r = ob.count_ops(params, scene, sampling, idx)
    wall = time.time() - t0
    flop_ro = sum(r["ops_rollout"][k] for k in FLOP_OPS)
    flop_sc = sum(r["ops_scoring"][k] for k in FLOP_OPS)
    n = len(idx)
    fis = params.fis.force_factor > 0
    w_survey = survey_w(cfg, fis)
    # per candidate-step: what the reference executes for ONE candidate and ONE step of the horizon, averaged over the sample.
    # Rejected candidates (the generator stops at the first step that violates the velocity limits) roll fewer steps; the
    # rollout figure divides by the steps actually rolled, the scoring figure by T x the candidates that reached the critics.
    per_step_rollout = flop_ro / max(1, r["steps_rolled"])
    per_step_scoring = flop_sc / max(1, r["n_generated"] * T)
    return {
        "config": cfg_name, "seed": seed, "candidates": int(C), "steps": int(T), "sampled_candidates": int(n),
        "generated": int(r["n_generated"]), "candidate_steps_rolled": int(r["steps_rolled"]),
        "objects": {"static": cfg.n_obstacles, "people": cfg.n_people, "groups": cfg.n_groups}, "fis": bool(fis),
        "ops_rollout": r["ops_rollout"], "ops_scoring": r["ops_scoring"],
        "flop_per_candidate_step_rollout": per_step_rollout, "flop_per_candidate_step_scoring": per_step_scoring,
        "flop_per_candidate_step": per_step_rollout + per_step_scoring,
        # what one planning cycle of the reference executes, extrapolated from the sample (rejections included as sampled)
        "flop_per_cycle": (flop_ro + flop_sc) * (C / n),
        "survey_estimate_W": w_survey, "instrumented_over_survey": (per_step_rollout + per_step_scoring) / w_survey,
        "wall_s": round(wall, 2),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_flop_count.json"))
    ap.add_argument("--sample", type=int, default=512)
    ap.add_argument("--seeds", type=int, nargs="*", default=[0, 1, 2])
    a = ap.parse_args()
    rows = []
    for name in ("cfg0", "cfg1", "cfg2", "cfg3"):
        for seed in a.seeds:
            row = count(name, seed, a.sample)
            rows.append(row)
            print(f"{name} seed {seed}: {row['flop_per_candidate_step']:.0f} flop per candidate-step (rollout "
                  f"{row['flop_per_candidate_step_rollout']:.0f} + scoring {row['flop_per_candidate_step_scoring']:.0f}); survey W = "
                  f"{row['survey_estimate_W']}; {row['wall_s']} s", flush=True)
    by_cfg = {}
    for name in ("cfg0", "cfg1", "cfg2", "cfg3"):
        rs = [r for r in rows if r["config"] == name]
        by_cfg[name] = {"flop_per_candidate_step": float(np.median([r["flop_per_candidate_step"] for r in rs])),
                        "flop_per_cycle": float(np.median([r["flop_per_cycle"] for r in rs])),
                        "survey_estimate_W": rs[0]["survey_estimate_W"], "seeds": [r["seed"] for r in rs]}
    out = {"what": "instrumented FP operation count of the CPU oracle (oracle/hmp_oracle_count.cpp: counting scalar, bit-identical "
                   "totals); flop = add + mul + div + sqrt + exp/pow + sin/cos/acos + atan2, one each; cmp / rnd listed only",
           "tool": "tools/count_flops.py", "median_over_seeds": by_cfg, "rows": rows}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
