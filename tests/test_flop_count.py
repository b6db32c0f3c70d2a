"""The instrumented operation count (oracle/hmp_oracle_count.cpp: the oracle's source text over a counting scalar) must be the
same computation as the oracle -- bit-identical totals -- and the committed profiles/r02_flop_count.json must be what the
tool produces (bench.py reads that file for roofline.algorithmic)."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
from humap_local_planner_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOP_OPS = ("add", "mul", "div", "sqrt", "exp", "trig", "atan2")


@pytest.mark.parametrize("cfg_name,seed", [("cfg0", 0), ("cfg0", 3), ("cfg1", 1)])
def test_counted_oracle_is_bit_identical(cfg_name, seed):
    cfg = scenes.CONFIGS[cfg_name]
    scene = scenes.make_scene(cfg, seed)
    params = scenes.make_params(cfg)
    sampling = scenes.make_sampling(cfg)
    C = ob.num_candidates(sampling)
    idx = np.unique(np.linspace(0, C - 1, min(C, 96)).astype(np.int64))
    r = ob.count_ops(params, scene, sampling, idx)
    want = np.array([ob.plan(params, scene, sampling, cand_range=(int(i), int(i) + 1), want=("totals",))["totals"][int(i)] for i in idx])
    assert np.array_equal(r["totals"], want, equal_nan=True)
    T = ob.num_steps(params, scene.world)
    assert r["n_generated"] == int((want != -1.0).sum())
    assert r["n_generated"] * T <= r["steps_rolled"] <= len(idx) * T
    # every category of the rollout is exercised; a step costs at least the static-object loop (> 40 flop per object)
    assert all(r["ops_rollout"][k] > 0 for k in FLOP_OPS)
    per_step = sum(r["ops_rollout"][k] for k in FLOP_OPS) / r["steps_rolled"]
    assert per_step > 40 * cfg.n_obstacles


def test_committed_flop_count_matches_the_tool():
    fc = json.load(open(os.path.join(ROOT, "profiles", "r02_flop_count.json")))
    row = next(r for r in fc["rows"] if r["config"] == "cfg0" and r["seed"] == 0)
    cfg = scenes.CONFIGS["cfg0"]
    scene = scenes.make_scene(cfg, 0)
    params = scenes.make_params(cfg)
    sampling = scenes.make_sampling(cfg)
    C = ob.num_candidates(sampling)
    r = ob.count_ops(params, scene, sampling, np.arange(C))
    assert row["sampled_candidates"] == C
    assert r["ops_rollout"] == row["ops_rollout"] and r["ops_scoring"] == row["ops_scoring"]
    # the survey's W overestimates the executed work of the reference formulation on every benchmark configuration
    for name, m in fc["median_over_seeds"].items():
        assert 0.3 < m["flop_per_candidate_step"] / m["survey_estimate_W"] < 1.0, name
