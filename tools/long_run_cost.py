"""Developer tool: cost of a large world as it ages.  Runs `workload` (c2 | c4 | c5 sizes of bench.py) for
`steps` steps in chunks and prints, per chunk, the mean step time, the iterations of the last step and whether
every step was proven identical to the sequential sweep.  Usage: long_run_cost.py c2 1000 100"""
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from crux_b200 import scenes, gpu

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
budget_s = float(sys.argv[4]) if len(sys.argv) > 4 else 120.0
n, L = {"c2": (65536, 64.0), "c4": (4194304, 256.0)}[wl]
pos, vel = scenes.sphere_worlds(1, n, L)
st = scenes.plane_bodies(scenes.box_planes(L))
g = gpu.GpuLargeWorld(n, len(st))
g.set(st, pos[0], vel[0])
if os.environ.get("PG_LARGE_TIMING"):
    g.set_timing(True)
t_start = time.time()
done = 0
while done < steps and time.time() - t_start < budget_s:
    ms, proven, worst = 0.0, True, 0.0
    for _ in range(chunk):
        if time.time() - t_start > budget_s:
            print("time budget reached inside the chunk", flush=True)
            break
        g.step(scenes.DT_60HZ, 1)
        m = g.last_step_ms()
        ms += m
        worst = max(worst, m)
        ex = g.exactness()
        proven = proven and ex["proven_exact"]
    done += chunk
    print(f"steps {done - chunk:5d}-{done:5d}: mean {ms / chunk:8.3f} ms/step, worst {worst:9.3f} ms, last step: {ex['iterations']} iterations, "
          f"{ex['candidates']} candidates, {ex['hits']} hits, all proven exact: {proven}", flush=True)
    if os.environ.get("PG_LARGE_TIMING"):
        print("      last step per kernel:", {k: round(v, 3) for k, v in g.kernel_ms().items()}, flush=True)
g.close()
