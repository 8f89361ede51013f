"""Developer tool: ms/step of every 10-step launch of the C3 batch over 1000 steps (run on a B200:
`gpurun -- python tools/per_launch_curve.py`).  Shows how the cost follows the worlds as they thermalise."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
w.set_spheres(0, st, pos, vel)
for c in range(100):
    w.step(scenes.DT_60HZ, 10)
    w.sync()
    print("ms/step", w.last_step_ms()/10)
