"""Developer tool: wall and CUDA-event time of pg_worlds_step_host / pg_worlds_step_host_packed (host buffers in,
host buffers out), with and without the physics kernel (n_steps = 0 leaves only the copy pipeline).
PG_HOST_CHUNKS / PG_HOST_DMA_QUEUES / PG_HOST_EVEN_CHUNKS / PG_HOST_NO_GRAPH select the pipeline's shape."""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("PG_HOST"))
for layout in ("split", "packed"):
    w = gpu.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
    w.set_spheres(0, st, pos, vel)
    if layout == "split":
        P = gpu.PinnedArray((nw, n, 3), np.float32); V = gpu.PinnedArray((nw, n, 3), np.float32); R = gpu.PinnedArray((nw, n), np.uint8)
        P.array[...] = pos; V.array[...] = vel; R.array[...] = 0
        call = lambda steps: w.step_host(scenes.DT_60HZ, steps, P.array, V.array, R.array)
    else:
        S = gpu.PinnedArray((nw,), w.packed_dtype())
        S.array["pos"] = pos; S.array["vel"] = vel; S.array["at_rest"] = 0
        call = lambda steps: w.step_host_packed(scenes.DT_60HZ, steps, S.array)
    for steps in (0, 1):
        for _ in range(3): call(steps)
        t0 = time.perf_counter()
        ev = 0.0
        for _ in range(30):
            call(steps); ev += w.last_step_ms()
        wall = (time.perf_counter() - t0) / 30 * 1e3
        print(f"[{tag}] {layout} n_steps={steps}: wall {wall:.3f} ms/call, events {ev/30:.3f} ms/call", flush=True)
    w.close()
