import sys, time
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
n, L = 4194304, 256.0
pos, vel = scenes.sphere_worlds(1, n, L)
st = scenes.plane_bodies(scenes.box_planes(L))
g = gpu.GpuLargeWorld(n, len(st))
g.set(st, pos[0], vel[0])
g.step(scenes.DT_60HZ, 3)
print("--- one step", flush=True)
g.step(scenes.DT_60HZ, 1)
print(g.last_step_ms(), g.exactness(), g.kernel_ms())
