import sys
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
w.set_spheres(0, st, pos, vel)
P = np.zeros((nw, n, 3), np.float32); V = np.zeros_like(P); R = np.zeros((nw, n), np.uint8)
w.step(scenes.DT_60HZ, 240); w.sync(); print("240 steps ms/step", w.last_step_ms()/240)
w.download_state(0, nw, P, V, R)
w.close()
def time_subset(idx, steps=5):
    m = len(idx)
    h = gpu.GpuWorlds(m, len(st), n, 0, max_events_per_world=0)
    h.set_spheres(0, st, P[idx], V[idx])
    h.upload_state(0, m, np.ascontiguousarray(P[idx]), np.ascontiguousarray(V[idx]), np.ascontiguousarray(R[idx]))
    h.step(scenes.DT_60HZ, steps); h.sync(); ms = h.last_step_ms()/steps
    h.close()
    return ms
groups = [np.arange(g*256, (g+1)*256) for g in range(16)]
tg = [time_subset(g) for g in groups]
print("group ms/step:", " ".join(f"{t:.2f}" for t in tg))
g = groups[int(np.argmax(tg))]
sub = [g[k*16:(k+1)*16] for k in range(16)]
ts = [time_subset(x) for x in sub]
print("subgroup ms/step:", " ".join(f"{t:.2f}" for t in ts))
s = sub[int(np.argmax(ts))]
tw = [time_subset(np.array([k])) for k in s]
print("world ms/step:", " ".join(f"{t:.2f}" for t in tw))
k = s[int(np.argmax(tw))]
print("slow world", k, "ms/step", max(tw))
np.savez("gpurun_out/slow_world.npz", P=P[k], V=V[k], R=R[k], k=k)
sp = np.linalg.norm(V[k], axis=1)
out = (np.abs(P[k][:,0])>5.2)|(np.abs(P[k][:,2])>5.2)|(P[k][:,1]<-0.2)|(P[k][:,1]>10.2)
print("max speed", sp.max(), "outside", np.where(out)[0], P[k][out], V[k][out])
