import subprocess, time, sys
exe="crux_b200/lib/crux_headless"; sc="tests/golden/scenes/bouncehouse.json"
def run(steps, extra):
    t=time.perf_counter(); r=subprocess.run([exe, sc, str(steps), "/tmp/o.bin", "--player"]+extra, capture_output=True, text=True); return time.perf_counter()-t, r
a,_=run(100,[]); b,r=run(5100,[])
print("per-frame mode: %.1f us per physics_step + scene update (15 bodies)"%((b-a)/5000*1e6), r.stdout.strip().splitlines()[-1], r.stderr.strip()[-200:])
a,_=run(100,["--device-steps"]); b,r=run(5100,["--device-steps"])
print("device-steps mode: %.1f us per frame"%((b-a)/5000*1e6), r.stdout.strip().splitlines()[-1], r.stderr.strip()[-200:])
