"""Developer tool: advance a large world to a hot state, then run two more steps (the target of an
`ncu -k regex:... --launch-skip N` capture)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from crux_b200 import scenes, gpu
skip = int(sys.argv[1]) if len(sys.argv) > 1 else 230
n, L = {"c2": (65536, 64.0), "c4": (4194304, 256.0), "m1": (1048576, 160.0)}[sys.argv[2] if len(sys.argv) > 2 else "c4"]
pos, vel = scenes.sphere_worlds(1, n, L)
st = scenes.plane_bodies(scenes.box_planes(L))
g = gpu.GpuLargeWorld(n, len(st))
g.set(st, pos[0], vel[0])
g.step(scenes.DT_60HZ, skip)
print("launches so far", g.launches, flush=True)
for s in range(2):
    g.step(scenes.DT_60HZ, 1)
    print("step", skip + s, "ms", round(g.last_step_ms(), 3), g.exactness(), flush=True)
