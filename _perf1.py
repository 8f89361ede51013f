import sys, time
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
w.set_spheres(0, st, pos, vel)
for steps in (1, 1, 10, 10, 100, 100):
    w.step(scenes.DT_60HZ, steps); w.sync()
    ms = w.last_step_ms()
    print(f"steps={steps} ms={ms:.3f} ms/step={ms/steps:.3f} body-steps/s={nw*n*steps/(ms*1e-3):.3e}")
print(w.stats())
