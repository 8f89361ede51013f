import sys, time
sys.path.insert(0, ".")
import numpy as np
from crux_b200 import scenes, gpu
nw, n = 4096, 256
pos, vel = scenes.sphere_worlds(nw, n, 10.0)
st = scenes.plane_bodies(scenes.box_planes(10.0))
w = gpu.GpuWorlds(nw, len(st), n, 0)
w.set_spheres(0, st, pos, vel)
hp = gpu.PinnedArray((nw, n, 3), np.float32); hv = gpu.PinnedArray((nw, n, 3), np.float32); hr = gpu.PinnedArray((nw, n), np.uint8)
w.download_state(0, nw, hp.array, hv.array, hr.array)
def t(f, reps=20):
    f(); t0=time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter()-t0)/reps*1e3
print("step_host n_steps=0 (copies+pack/unpack only): %.3f ms" % t(lambda: w.step_host(scenes.DT_60HZ, 0, hp.array, hv.array, hr.array)))
print("step_host n_steps=1: %.3f ms" % t(lambda: w.step_host(scenes.DT_60HZ, 1, hp.array, hv.array, hr.array)))
print("upload only: %.3f ms" % t(lambda: (w.upload_state(0, nw, hp.array, hv.array, hr.array), w.sync())))
print("download only: %.3f ms" % t(lambda: w.download_state(0, nw, hp.array, hv.array, hr.array)))
print("step only: %.3f ms" % t(lambda: (w.step(scenes.DT_60HZ, 1), w.sync())))
# pageable comparison
P = hp.array.copy(); V = hv.array.copy(); R = hr.array.copy()
print("step_host pageable n_steps=1: %.3f ms" % t(lambda: w.step_host(scenes.DT_60HZ, 1, P, V, R), 5))
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv,noheader", shell=True, capture_output=True, text=True).stdout)
