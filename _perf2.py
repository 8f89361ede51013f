import sys
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
w.set_spheres(0, st, pos, vel)
for steps in (4, 4, 4):
    w.step(scenes.DT_60HZ, steps); w.sync()
    print(steps, w.last_step_ms())
