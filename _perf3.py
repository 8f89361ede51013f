import sys, subprocess, threading, time
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
def fresh():
    w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
    w.set_spheres(0, st, pos, vel)
    return w
w = fresh()
w.step(scenes.DT_60HZ, 2); w.sync()
w.close(); w = fresh()
out=[]
for c in range(30):
    w.step(scenes.DT_60HZ, 10); w.sync(); out.append(w.last_step_ms()/10)
print("10-step chunks ms/step:", " ".join(f"{x:.3f}" for x in out))
w.close(); w = fresh()
w.step(scenes.DT_60HZ, 300); w.sync(); print("single 300-step launch ms/step:", w.last_step_ms()/300)
w.close(); w = fresh()
out=[]
for c in range(6):
    w.step(scenes.DT_60HZ, 1); w.sync(); out.append(w.last_step_ms())
print("1-step launches ms:", " ".join(f"{x:.3f}" for x in out))
r = subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader", shell=True, capture_output=True, text=True); print(r.stdout.strip())
