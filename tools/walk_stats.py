"""Developer tool: interval-walk node counters.  Needs a debug library built with -DPG_WALK_STATS
(`nvcc ... -DPG_WALK_STATS` on crux_b200/csrc/*.cu, linked to crux_b200/lib/alt/libphysics_gpu_ws.so);
prints calls / nodes / longest walk per kind (sphere-sphere, sphere-plane, box, leaf scans) per launch."""
import os, sys, ctypes as C
os.environ["CRUX_PG_LIB"] = os.path.abspath("crux_b200/lib/alt/libphysics_gpu_ws.so")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from crux_b200 import scenes, gpu
L = C.CDLL(os.environ["CRUX_PG_LIB"])
def stats(reset=1):
    a = (C.c_ulonglong * 16)(); L.pg_debug_walk_stats(a, reset); return np.array(a).reshape(4, 4)
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
w.set_spheres(0, st, pos, vel)
for c in range(100):
    w.step(scenes.DT_60HZ, 10); w.sync()
    s = stats()
    if c % 10 == 9 or c > 95:
        print(c, "ms/step", round(w.last_step_ms()/10, 3), "ss calls/nodes/max/heavy", s[0].tolist(), "sp", s[1].tolist(), "scan calls/cands/max", s[3].tolist())
