#!/usr/bin/env python
"""Benchmark of the trajectory sampling + scoring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cfg cfg2]

A "step" is one planning cycle: every candidate of the workload rolled out, scored by all critics and the
argmin selected. The default workload is BASELINE.json configs[2] ("crowd stress": 65 536 candidates x
50 people x 8 F-formation groups x 500 obstacle points, 5 s horizon) -- the configuration the metric's
"p50 planning-cycle latency at 64k candidates" is quoted on; it fits one GPU.

  value  candidates/s with the scene already resident in HBM (hmp_replan_resident), timed with CUDA events
         on the library's launching stream (HmpResult.gpu_ms), max over ranks.
  e2e    the same metric through the public C-ABI calls a planner makes every cycle, host buffers in, host
         result out: hmp_set_costmap + 4 x hmp_set_mapgrid + hmp_set_footprint + hmp_plan (wall clock).
  N > 1  independent scenes, one per rank (scene seed = rank), no collective on the data path: weak scaling.

`--impl reference` times the reference's CPU implementation of the path (oracle/_ref: its own sources compiled in
place; falls back to the oracle restatement if that library is absent), FP64, all host threads, on a bounded candidate
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "candidate trajectories scored/sec"
UNIT = "candidates/s"


def algorithmic_flops_per_candidate(cfg, params, n_static, n_dynamic, fis_on, T, K=8, V=16, c_cells=45):
    """W = 294 + 86 S + 103 D + 185 P + 45 G + (K+1)(16 V + 6 c) [+ 2900 D with FIS] per step."""
    P, G = cfg.n_people, cfg.n_groups
    W = 294 + 86 * n_static + 103 * n_dynamic + 185 * P + 45 * G + (K + 1) * (16 * V + 6 * c_cells)
    if fis_on:
        W += 2900 * n_dynamic
    return T * W + 60, W


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        smax = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if "Active" in r[3 + i] and "Not" not in r[3 + i]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(cfg, scene, params, sampling, seconds_budget=12.0):
    """The reference path on the host cores, bounded sample, all host threads.

    kind "reference": oracle/_ref/libhmp_ref.so -- the reference's own first-party sources for the path (generator, SFM,
    World, FIS, conductor, critics) compiled in place against the stand-in third-party headers of oracle/ref_shim/ --
    when that library was built (it travels to the GPU box with the snapshot); kind "port": the oracle restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    import concurrent.futures as cf
    cores = os.cpu_count() or 1
    impl = "ref" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhmp_ref.so")) else "oracle"
    kind = "reference" if impl == "ref" else "port"
    Cn = ob.num_candidates(sampling)
    # calibrate on a few candidates, then size the sample for ~seconds_budget of wall time
    probe = np.unique(np.linspace(0, Cn - 1, 4).astype(int))
    t0 = time.perf_counter()
    for i in probe:
        ob.plan(params, scene, sampling, cand_range=(int(i), int(i) + 1), want=("totals",), impl=impl)
    per = (time.perf_counter() - t0) / len(probe)
    n = int(min(Cn, max(cores, seconds_budget / per * cores)))
    idx = np.unique(np.linspace(0, Cn - 1, n).astype(int))
    chunks = np.array_split(idx, cores * 2)

    def work(ch):
        for i in ch:
            ob.plan(params, scene, sampling, cand_range=(int(i), int(i) + 1), want=("totals",), impl=impl)
        return len(ch)

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(cores) as ex:
        done = sum(ex.map(work, chunks))
    dt = time.perf_counter() - t0
    what = ("reference first-party sources compiled in place (oracle/_ref, stand-in third-party headers)" if impl == "ref"
            else "oracle restatement (oracle/hmp_oracle.cpp)")
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{done} of {Cn} candidates of {cfg.name} (evenly spaced over the sampling grid), {dt:.1f} s wall, "
                      f"single-thread rate {1.0 / per:.1f} candidates/s; {what}"}, dt, done


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path (oracle/_ref, else the oracle port), bounded sample per step, rank 0 only."""
    if rank != 0:
        return
    from humap_local_planner_b200 import scenes
    cfg = scenes.CONFIGS[args.cfg]
    scene = scenes.make_scene(cfg, 0)
    params = scenes.make_params(cfg, fis=bool(args.fis))
    sampling = scenes.make_sampling(cfg)
    per_step = max(3.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    times, counts = [], []
    base = None
    for s in range(args.warmup + args.steps):
        base, dt, done = cpu_baseline(cfg, scene, params, sampling, seconds_budget=per_step)
        if s >= args.warmup:
            times.append(dt)
            counts.append(done)
    value = sum(counts) / sum(times)
    base["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_name(cfg, args), "candidates_per_step_sample": int(np.mean(counts))},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _workload_name(cfg, args):
    return (f"{cfg.name}: {cfg.n_people} people, {cfg.n_groups} F-formation groups, {cfg.n_obstacles} obstacle points, "
            f"{cfg.size}x{cfg.size} costmap @ {cfg.resolution} m, horizon {cfg.sim_time} s @ {cfg.sim_granularity} s, "
            f"FIS {'on' if args.fis else 'off'}")


def _make_scene_job(a):
    from humap_local_planner_b200 import scenes
    name, seed = a
    sc = scenes.make_scene(scenes.CONFIGS[name], seed)
    w = sc.world
    flat = dict(world=[w.robot_x, w.robot_y, w.robot_yaw, w.vel_x, w.vel_y, w.vel_th, w.goal_local_x, w.goal_local_y,
                       w.goal_local_yaw, w.goal_x, w.goal_y, w.goal_yaw],
                obstacles=bytes(memoryview(sc._obstacles))[: w.n_obstacles * 80], n_obstacles=w.n_obstacles,
                people=bytes(memoryview(sc._people))[: w.n_people * 80], n_people=w.n_people,
                groups=bytes(memoryview(sc._groups))[: w.n_groups * 64], n_groups=w.n_groups,
                cells=sc.cells, grids=[g.astype(np.float32) for g in sc.grids], hv=sc.hv_prev)
    return flat


def run_batched(args, rank, local_rank, world, barrier):
    """BASELINE config 4: `--scenes` independent worlds x cfg3's 4096 candidates each, scene s -> rank s mod N, one
    hmp_plan_batch launch per rank per step, host gather of the per-scene argmins (no collective on the data path).
    Total work is fixed, so this is STRONG scaling."""
    import ctypes as Cc
    import multiprocessing as mp
    import torch
    from humap_local_planner_b200 import Planner, scenes, capi
    from humap_local_planner_b200.sharding import scenes_for_rank, gather_scene_results
    cfg = scenes.CONFIGS[args.cfg]
    S = args.scenes
    mine = scenes_for_rank(S, rank, world)
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        flats = pool.map(_make_scene_job, [(args.cfg, s) for s in mine], chunksize=4)
    worlds, keep = [], []
    for f in flats:
        w = capi.HmpWorld()
        (w.robot_x, w.robot_y, w.robot_yaw, w.vel_x, w.vel_y, w.vel_th, w.goal_local_x, w.goal_local_y, w.goal_local_yaw,
         w.goal_x, w.goal_y, w.goal_yaw) = f["world"]
        ob = (capi.HmpObstacle * max(1, f["n_obstacles"])).from_buffer_copy(f["obstacles"].ljust(80 * max(1, f["n_obstacles"]), b"\0"))
        pe = (capi.HmpPerson * max(1, f["n_people"])).from_buffer_copy(f["people"].ljust(80 * max(1, f["n_people"]), b"\0"))
        gr = (capi.HmpGroup * max(1, f["n_groups"])).from_buffer_copy(f["groups"].ljust(64 * max(1, f["n_groups"]), b"\0"))
        w.obstacles, w.people, w.groups = ob, pe, gr
        w.n_obstacles, w.n_people, w.n_groups = f["n_obstacles"], f["n_people"], f["n_groups"]
        keep.append((ob, pe, gr))
        worlds.append(w)
    cells = np.stack([f["cells"] for f in flats])
    grids = [np.stack([f["grids"][g] for f in flats]).astype(np.float64) for g in range(4)]
    hv = np.array([f["hv"] for f in flats])
    first = scenes.make_scene(cfg, mine[0])
    params = scenes.make_params(cfg, fis=bool(args.fis))
    sampling = scenes.make_sampling(cfg)
    pl = Planner(local_rank)
    pl.set_precision(int(args.precise))
    pl.set_sweep_layout(int(args.layout))
    pl.set_params(params)
    pl.set_scene(first)
    n_local = len(mine)
    res = pl.plan_batch(worlds, cells, grids, sampling, hv_prev=hv)
    C = res[0].n_candidates
    for _ in range(max(1, args.warmup - 1)):
        pl.replan_resident(n_local)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pl.launch_count()
    dev_ms = []
    for _ in range(args.steps):
        torch.cuda.synchronize()
        dev_ms.append(pl.replan_resident(n_local)[0].gpu_ms)   # per-step inputs (n_local x 680 kB) exceed the 126 MB L2
    barrier()
    launches = pl.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = torch.tensor([sum(dev_ms) * 1e-3], dtype=torch.float64, device="cuda")
    e2e_t = []
    for _ in range(min(args.steps, 3)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pl.plan_batch(worlds, cells, grids, sampling, hv_prev=hv)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    t_e2e = torch.tensor([statistics.mean(e2e_t)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    # host gather of the per-scene argmins (16 bytes per scene)
    gathered = gather_scene_results({s: (r.best_index, r.best_total) for s, r in zip(mine, res)}, S)
    if rank == 0:
        n_cells = cells[0].size
        line = {
            "metric": METRIC, "value": S * C * args.steps / float(t_dev.item()), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(t_dev.item()) / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {0: "f32", 1: "f64", 2: "f32 sweep + f64 refinement of the leaders"}[int(args.precise)],
            "data": "synthetic",
            "config": {"workload": f"batched scenes: {S} independent worlds x {C} candidates ({_workload_name(cfg, args)})",
                       "parallelism": f"scene s -> rank s mod {world}, one hmp_plan_batch launch per rank, host gather of argmins, no collective",
                       "l2": f"inputs per step ({n_local} scenes x {(n_cells * 17) // 1024} kB) exceed L2", "timing": "CUDA events on the launching stream"},
            "e2e": {"value": S * C / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": int(n_local * n_cells * 17),
                    "d2h_bytes_per_step": int(n_local * Cc.sizeof(capi.HmpResult))},
            "gpu_launches": int(launches), "clocks": clocks,
            "scenes_with_valid_winner": int(sum(1 for b, _ in gathered if b >= 0)),
        }
        print(json.dumps(line))
    pl.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", default="cfg2")
    ap.add_argument("--fis", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scenes", type=int, default=0,
                    help="batched-scenes mode (BASELINE config 4, use with --cfg cfg3): this many independent worlds in total, "
                         "sharded over the ranks, one hmp_plan_batch launch per rank per step; 0 = single-scene mode")
    ap.add_argument("--precise", type=int, default=2, help="hmp_set_precision mode: 0 FP32 object loops, 1 FP64 (exact-parity mode), 2 FP32 sweep + FP64 refinement of the leaders")
    ap.add_argument("--layout", type=int, default=0, help="hmp_set_sweep_layout: 0 automatic, 1 one warp per candidate, 2 one thread per candidate")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from humap_local_planner_b200 import Planner, scenes
    from humap_local_planner_b200.sharding import scenes_for_rank

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.scenes > 0:
        run_batched(args, rank, local_rank, world, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    cfg = scenes.CONFIGS[args.cfg]
    # independent scenes are sharded over ranks (scene s -> rank s mod N), one scene per rank per step: weak scaling. Every
    # rank plans its own copy of the SAME synthetic world (seed 0), so that the work per GPU is identical at every N (the
    # cost of a cycle varies by ~10 % between seeds, which would otherwise show up as a scaling loss through the max over ranks)
    my_scene_ids = scenes_for_rank(world, rank, world)
    assert my_scene_ids == [rank]
    scene = scenes.make_scene(cfg, seed=0)
    params = scenes.make_params(cfg, fis=bool(args.fis))
    sampling = scenes.make_sampling(cfg)
    pl = Planner(local_rank)
    pl.set_precision(int(args.precise))
    pl.set_sweep_layout(int(args.layout))
    pl.set_params(params)

    def full_cycle():
        # what HumapPlanner does every control cycle before and at the seam (humap_planner.cpp:1054-1141, :1294-1377)
        pl.set_costmap(scene.cells, scene.origin_x, scene.origin_y, scene.resolution)
        for g in range(4):
            pl.set_mapgrid(g, scene.grids[g], scene.hv_prev[g])
        pl.set_footprint(scene.footprint)
        return pl.plan(scene.world, sampling, want_poses=True)

    res, _ = full_cycle()
    C = res.n_candidates
    T = pl.num_steps()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- value: inputs resident in HBM, CUDA-event time of the kernels on the launching stream ------------------
    for _ in range(args.warmup):
        flush.zero_()
        pl.replan_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pl.launch_count()
    dev_ms, sel_ms = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()          # L2 flush between timed iterations (not inside the event-timed region)
        torch.cuda.synchronize()
        r = pl.replan_resident()[0]
        dev_ms.append(r.gpu_ms)
        sel_ms.append(r.gpu_ms_select)
    barrier()
    wall_resident = time.perf_counter() - t_wall0
    launches = pl.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = torch.tensor([sum(dev_ms) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev = float(t_dev.item())

    # ---- e2e: host buffers through the public C-ABI, every step uploads the cycle's inputs and reads the result ----
    for _ in range(2):
        full_cycle()
    barrier()
    e2e_times = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r, poses = full_cycle()
        e2e_times.append(time.perf_counter() - t0)
    barrier()
    # variant: MapGridCostFunction::prepare() replaced as well (wave fronts computed on the device from the plan poses)
    def full_cycle_device_grids():
        pl.set_costmap(scene.cells, scene.origin_x, scene.origin_y, scene.resolution)
        for g, (plan, local_goal) in enumerate(scene.plans):
            pl.compute_mapgrid(g, plan, local_goal, scene.hv_prev[g])
        pl.set_footprint(scene.footprint)
        return pl.plan(scene.world, sampling, want_poses=True)
    full_cycle_device_grids()
    e2e_dev_times = []
    for _ in range(min(args.steps, 10)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r_dev, _ = full_cycle_device_grids()
        e2e_dev_times.append(time.perf_counter() - t0)
    assert r_dev.best_index == r.best_index and r_dev.best_total == r.best_total, "device wave front changed the selection"
    barrier()
    t_e2e = torch.tensor([sum(e2e_times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())

    if rank == 0:
        import ctypes as Cc
        from humap_local_planner_b200 import capi
        n_cells = scene.cells.size
        h2d = n_cells + 4 * n_cells * 4 + Cc.sizeof(capi.HmpParams) + scene.world.n_obstacles * 64 + scene.world.n_people * 64 + \
            scene.world.n_groups * 32 + 10 * 64 * 8
        d2h = (14 + 3 + 3 * T + 1) * 8 + 8 + 64
        nd = sum(1 for i in range(scene.world.n_obstacles) if scene.world.obstacles[i].force_dynamic or
                 (scene.world.obstacles[i].vx ** 2 + scene.world.obstacles[i].vy ** 2) ** 0.5 > 0.035)
        ns = scene.world.n_obstacles - nd
        flops_cand, W = algorithmic_flops_per_candidate(cfg, params, ns, nd, bool(args.fis), T)
        sel_s = statistics.mean(sel_ms) * 1e-3
        achieved = C * flops_cand / sel_s / 1e12
        mode = pl.last_sweep_mode()
        sweep_name = (f"sweep_tpc_kernel (one thread per candidate, {mode} threads per block)" if mode
                      else "plan_kernel<false,float> (one warp per candidate)")
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # DRAM bytes per launch of the dominant kernel from the latest committed ncu --set full capture
            import glob
            tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))[-1]
            tj = json.load(open(tf))
            if args.cfg == "cfg2":
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak_fp32 = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        peak_fp32_measured = pl.measure_fp32_peak()       # FFMA micro-benchmark on this GPU, same launch shape as the sweep
        hbm_peak = float(peaks.get("hbm_gbs", 0.0)) or None
        line = {
            "metric": METRIC, "value": world * C * args.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {0: "f32", 1: "f64", 2: "f32 sweep + f64 refinement of the leaders"}[int(args.precise)], "data": "synthetic",
            "config": {"workload": _workload_name(cfg, args), "candidates": C, "steps_per_rollout": T,
                       "parallelism": f"independent scenes x{world} (copies of the seed-0 world), one per rank, no collective",
                       "l2": "flushed with a 256 MiB write between timed iterations", "timing": "CUDA events on the launching stream"},
            "p50_cycle_ms": statistics.median(dev_ms), "p99_cycle_ms": sorted(dev_ms)[min(len(dev_ms) - 1, int(0.99 * len(dev_ms)))],
            "p50_cycle_ms_e2e": 1e3 * statistics.median(e2e_times),
            "p50_cycle_ms_e2e_device_mapgrids": 1e3 * statistics.median(e2e_dev_times),
            "wall_ms_per_step_resident": 1e3 * wall_resident / args.steps,
            "e2e": {"value": world * C * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp32-cuda-core", "achieved": achieved, "peak": peak_fp32, "unit": "TFLOP/s",
                         "frac": achieved / peak_fp32, "traffic": traffic,
                         "peak_measured": peak_fp32_measured, "frac_of_measured": achieved / peak_fp32_measured,
                         "hbm": {"achieved_gbs": (traffic / sel_s / 1e9) if traffic else None, "peak_gbs": hbm_peak},
                         "note": f"algorithmic flop per candidate-step W={W} (SURVEY.md 8d formula), per candidate T*W+60; dominant kernel "
                                 f"{sweep_name} avg {1e3 * sel_s:.3f} ms; peak = nominal 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz "
                                 "(MEASURED_PEAKS.json has no FP32 CUDA-core entry; the path is not HBM- or tensor-bound)"},
            "best_index": int(r.best_index), "best_total": float(r.best_total), "n_valid": int(r.n_valid),
        }
        if not args.no_cpu_baseline:
            base, _, _ = cpu_baseline(cfg, scene, params, sampling, seconds_budget=12.0)
            line["cpu_baseline"] = base
        print(json.dumps(line))
    pl.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
