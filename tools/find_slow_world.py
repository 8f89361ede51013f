"""Developer tool: at a given launch of the 1000-step C3 run, bisect the batch for the slowest world and
save its state to gpurun_out/slow_world2.npz.  This is how the pathological worlds of round 1 were found
(resting body pushed along a plane, grazing pairs, a sphere that tunnelled out of the box)."""
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
P = np.zeros((nw, n, 3), np.float32); V = np.zeros_like(P); R = np.zeros((nw, n), np.uint8)
out = []
for c in range(100):
    w.step(scenes.DT_60HZ, 10); w.sync(); out.append(w.last_step_ms()/10)
print("ms/step per 10-step launch:", " ".join(f"{x:.2f}" for x in out))
worst = 99
# replay to just before the worst launch
w.close(); w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0); w.set_spheres(0, st, pos, vel)
w.step(scenes.DT_60HZ, worst * 10); w.sync()
w.download_state(0, nw, P, V, R)
w.close()
def time_subset(idx, steps=10):
    m = len(idx)
    h = gpu.GpuWorlds(m, len(st), n, 0, max_events_per_world=0)
    h.set_spheres(0, st, P[idx], V[idx])
    h.upload_state(0, m, np.ascontiguousarray(P[idx]), np.ascontiguousarray(V[idx]), np.ascontiguousarray(R[idx]))
    h.step(scenes.DT_60HZ, steps); h.sync(); ms = h.last_step_ms()/steps
    h.close()
    return ms
groups = [np.arange(g*256, (g+1)*256) for g in range(16)]
tg = [time_subset(g) for g in groups]
print("worst launch", worst, "group ms/step:", " ".join(f"{t:.2f}" for t in tg))
g = groups[int(np.argmax(tg))]
sub = [g[k*16:(k+1)*16] for k in range(16)]
ts = [time_subset(x) for x in sub]
s = sub[int(np.argmax(ts))]
tw = [time_subset(np.array([k])) for k in s]
print("world ms/step:", " ".join(f"{t:.2f}" for t in tw))
k = s[int(np.argmax(tw))]
print("slow world", k, "ms/step", max(tw))
np.savez("gpurun_out/slow_world2.npz", P=P[k], V=V[k], R=R[k], k=k)
