import sys
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=4096)
w.set_spheres(0, st, pos, vel)
P = np.zeros((nw, n, 3), np.float32); V = np.zeros_like(P); R = np.zeros((nw, n), np.uint8)
for c in range(30):
    w.step(scenes.DT_60HZ, 10); w.sync(); ms = w.last_step_ms()/10
    ev = w.events()
    per_world = np.bincount(ev["world"], minlength=nw)
    w.download_state(0, nw, P, V, R)
    sp = np.linalg.norm(V, axis=2)
    out = (np.abs(P[...,0])>5.2)|(np.abs(P[...,2])>5.2)|(P[...,1]<-0.2)|(P[...,1]>10.2)
    print(f"chunk {c:2d} ms/step {ms:.3f} events/world-step mean {per_world.mean()/10:.1f} max {per_world.max()/10:.1f} (world {per_world.argmax()}) overflow {w.events_overflowed} maxspeed {np.nanmax(sp):.1f} nonfinite {np.isnan(P).sum()} outside {out.sum()} rest {R.sum()}")
top = np.argsort(-per_world)[:5]
print("top worlds", top, per_world[top])
k = top[0]
print("world", k, "max speed", sp[k].max(), "outside", out[k].sum(), "rest", R[k].sum())
np.save("gpurun_out/bad_world_P.npy", P[k]); np.save("gpurun_out/bad_world_V.npy", V[k]); np.save("gpurun_out/bad_world_R.npy", R[k])
print("bad world index", k)
