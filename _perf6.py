import sys, time
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
for name, n, L, steps in (("c2", 65536, 64.0, 50), ("c4", 4194304, 256.0, 10)):
    pos, vel = scenes.sphere_worlds(1, n, L)
    st = scenes.plane_bodies(scenes.box_planes(L))
    g = gpu.GpuLargeWorld(n, len(st))
    g.set(st, pos[0], vel[0])
    g.step(scenes.DT_60HZ, 3)
    t0 = time.perf_counter(); g.step(scenes.DT_60HZ, steps); wall = time.perf_counter() - t0
    ms = g.last_step_ms()
    print(name, f"n={n} steps={steps} ms/step={ms/steps:.3f} wall ms/step={wall/steps*1e3:.3f} body-steps/s={n*steps/(ms*1e-3):.3e}", g.exactness())
    km = g.kernel_ms(); tot = sum(km.values())
    print("  kernel ms per step:", {k: round(v/steps, 4) for k, v in km.items()}, "sum", round(tot/steps, 3))
    print("  stats", g.stats())
    g.close()
