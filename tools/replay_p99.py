import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np
from humap_local_planner_b200 import Planner, replay, config
pl = Planner(0)
replay.run_replay(pl, n_cycles=60)
rows = []
orig = pl.plan
def plan(*a, **k):
    r = orig(*a, **k)
    rows.append((r[0].gpu_ms, r[0].gpu_ms_select, pl.last_num_leaders(), pl.last_num_leaders_round2(), pl.last_fallback_rounds(), r[0].n_valid, r[0].n_generated))
    return r
pl.plan = plan
log = replay.run_replay(pl, n_cycles=1000, sampling_axes=config.SAMPLING_64K)
a = np.array(rows)
order = np.argsort(-a[:, 0])
print("p50 gpu", np.percentile(a[:, 0], 50), "p99", np.percentile(a[:, 0], 99))
print("slowest 12: gpu_ms, sweep_ms, leaders, leaders2, fallback, n_valid, n_generated")
for i in order[:12]: print(i, [round(float(x), 2) for x in a[i]])
print("median rows:"); 
for i in order[len(order)//2: len(order)//2 + 3]: print(i, [round(float(x), 2) for x in a[i]])
print("corr(sweep, n_generated)", np.corrcoef(a[:,1], a[:,6])[0,1], "fallback cycles", int((a[:,4]>0).sum()))
