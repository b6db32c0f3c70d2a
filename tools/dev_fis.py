import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob
from humap_local_planner_b200 import Planner
pl = Planner(0)
rng = np.random.default_rng(5)
N = 40000
x = rng.uniform(-np.pi, np.pi, (N, 4))
got = pl.debug_fis(x)
L = ob.lib(); L.orc_fis_process.argtypes = [C.c_double]*4 + [C.c_void_p]*3
ref = np.zeros((N, 2)); out = (C.c_double*3)()
for i in range(N):
    L.orc_fis_process(*x[i], out, None, None); ref[i] = (out[0], out[1])
dv = np.abs((got[:,0]-ref[:,0]+np.pi)%(2*np.pi)-np.pi); dm = np.abs(got[:,1]-ref[:,1])
for t in (1e-6,1e-5,1e-4,1e-3,1e-2):
    print(f"thr {t:g}: dv>{(dv>t).mean():.5f} dm>{(dm>t).mean():.5f}")
bad = np.argsort(-np.maximum(dv,dm))[:8]
for i in bad:
    print("in(deg)", np.round(np.degrees(x[i]),2), "gpu", got[i], "ref", ref[i])
