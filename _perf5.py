import sys
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
w.set_spheres(0, st, pos, vel)
w.step(scenes.DT_60HZ, 10); w.sync()
out=[]
for c in range(3):
    w.step(scenes.DT_60HZ, 100); w.sync(); out.append(w.last_step_ms()/100)
print("100-step launches ms/step:", " ".join(f"{x:.3f}" for x in out), " body-steps/s %.3e" % (nw*n/(out[-1]*1e-3)))
