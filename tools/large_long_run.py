"""Developer tool: run the C2 world (65,536 spheres; `large_long_run.py STEPS c4` for the 4 M one) for many steps, then print the per-iteration trace of a
few steps (PG_LARGE_DEBUG=1) and the per-kernel times.  Shows what a settled, piled-up world costs."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from crux_b200 import scenes, gpu
skip = int(sys.argv[1]) if len(sys.argv) > 1 else 900
n, L = {"c2": (65536, 64.0), "c4": (4194304, 256.0)}[sys.argv[2] if len(sys.argv) > 2 else "c2"]
pos, vel = scenes.sphere_worlds(1, n, L)
st = scenes.plane_bodies(scenes.box_planes(L))
g = gpu.GpuLargeWorld(n, len(st))
g.set(st, pos[0], vel[0])
g.step(scenes.DT_60HZ, skip)
P, V, R = g.download_state()
print("after", skip, "steps: at rest", int((R != 0).sum()), "y range", float(P[:, 1].min()), float(P[:, 1].max()),
      "max speed", float(np.linalg.norm(V, axis=1).max()), "bodies below the floor", int((P[:, 1] < -1).sum()))
os.environ["PG_LARGE_DEBUG"] = "1"
g.set_timing(True)
for s in range(2):
    g.step(scenes.DT_60HZ, 1)
    print("step", skip + s, "ms", round(g.last_step_ms(), 3), {k: round(v, 3) for k, v in g.kernel_ms().items()}, g.exactness())
