import sys, time, json, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
from humap_local_planner_b200 import Planner, scenes
import oracle_binding as ob
pl = Planner(0)
out = {}
for name in ("cfg0", "cfg2"):
    cfg = scenes.CONFIGS[name]
    for seed in (0, 1):
        sc = scenes.make_scene(cfg, seed)
        pl.set_params(scenes.make_params(cfg))
        pl.set_costmap(sc.cells, sc.origin_x, sc.origin_y, sc.resolution)
        plan = np.array([[0.1 * i, 0.02 * i] for i in range(30)])
        ts = []
        for it in range(30):
            t0 = time.perf_counter()
            for g, lg in enumerate((False, True, False, True)):
                pl.compute_mapgrid(g, plan, lg, 0.0)
            grids = pl.get_mapgrid(0, sc.cells.shape)   # synchronises all four
            ts.append(1e3 * (time.perf_counter() - t0))
        ok = True
        for g, lg in enumerate((False, True, False, True)):
            dev = pl.get_mapgrid(g, sc.cells.shape)
            ref = scenes.mapgrid_wavefront(sc.cells, sc.origin_x, sc.origin_y, sc.resolution, plan, lg)
            ok &= bool(np.array_equal(dev, ref))
        out[f"{name}_s{seed}"] = {"four_wavefronts_plus_one_readback_ms_p50": float(np.median(ts)), "equal_to_host_bfs": ok}
print(json.dumps(out, indent=1))
