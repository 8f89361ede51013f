"""Developer tool: per-step timing of the C5 level world (1 M spheres): step, k_pass3, k_eval, event counts."""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from crux_b200 import scenes, gpu
n, L = 1048576, 160.0
statics, p0, v0, players = scenes.level_world(n, L, seed_world=0, with_player=False)
g = gpu.GpuLargeWorld(n, len(statics))
g.set(gpu.prepare(statics), p0, v0)
g.set_timing(True)
for s in range(24):
    g.step(scenes.DT_60HZ, 1)
    km = g.kernel_ms()
    ev = g.events()
    nb = int(((ev["a_class"] == 0) & (ev["b_class"] == 1)).sum()); npl = int(((ev["pass_"] == 3)).sum()) - nb
    print(s, f"step {g.last_step_ms():.3f} pass3 {km['k_pass3']:.3f} eval {km['k_eval']:.3f} box-events {nb} plane-events {npl} pass4 {int((ev['pass_']==4).sum())}", g.exactness()["iterations"])
