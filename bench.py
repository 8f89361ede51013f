#!/usr/bin/env python3
"""bench.py -- body-steps/s of the Crux physics step on B200 (and the reference CPU step beside it).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU physics

Workload (config.workload = "c3"): BASELINE.json configs[2], the one the headline metric is quoted
on -- 4096 independent 256-sphere "bouncehouse" worlds per GPU (6 wall planes, r = 0.15, LCG seed
12345 + world index), dt = 1/60 -> clamped to 0.01666.  One "step" is one
physics_step of every world.  Worlds shard across ranks with no data-path collective (weak scaling:
4096 worlds per GPU); the only exchange is the 8-double aggregate statistics vector, all-reduced
over NCCL once per launch.

Timing: W warm-up steps, then K timed steps issued as launches of --steps-per-launch steps.  Every
launch is bracketed by CUDA events on the library's own stream (pg_worlds_last_step_ms); L2 is
flushed between launches outside the event pairs.  ms_per_step = max over ranks of the summed event
times / K.  `e2e` repeats the measurement through the host-buffer API (pinned host arrays in,
pinned host arrays out, every step).  `cpu_baseline` times the reference CPU step on this box.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from crux_b200 import scenes  # noqa: E402

METRIC = "body_steps_per_sec"
UNIT = "body-steps/s"
ALGO_BYTES_PER_BODY_STEP = 52          # integrate R 28 B + W 24 B
WORKLOADS = {
    "c3": dict(n_worlds=4096, n_spheres=256, L=10.0, desc="4096 batched 256-sphere bouncehouse worlds per GPU"),
    "c3-small": dict(n_worlds=256, n_spheres=256, L=10.0, desc="256 batched 256-sphere worlds per GPU (smoke size)"),
    # single large worlds (uniform-grid path); under N > 1 every rank runs an independent replica
    "c2": dict(n_worlds=1, n_spheres=65536, L=64.0, desc="one world, 65,536 spheres in a closed box", large=True),
    "c4": dict(n_worlds=1, n_spheres=4194304, L=256.0, desc="one world, 4,194,304 spheres in a closed box", large=True),
    "c5": dict(n_worlds=1, n_spheres=1048576, L=160.0, large=True, level=True,
               desc="static AABB slab level (8-unit lattice) + 1,048,576 spheres + player capsule"),
}
# algorithmic bytes per body-step of the streaming kernels of the large-world path (restated for
# two phase entries per body -- DESIGN.md section 4.4): pass 3 reads the 32-byte state and writes the two entry radii;
# the counting sort reads the state and the radii and writes cell + rank per entry, then writes one 32-byte record per
# entry; the pair search reads both records of a body; the final kernel is the survey's "integrate"
LARGE_KERNEL_BYTES = {"k_pass3": 40, "k_cell_count": 56, "k_scan": 12, "k_scatter": 120, "k_pairs": 64, "k_finalize": 52}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference CPU step (oracle/_ref = the unmodified reference compiled here; else the plain-C port)
# ------------------------------------------------------------------------------------------------

def _cpu_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ora
    if os.path.exists(ora.REF_SO):
        return ora, ora.RefLib(), "reference"
    if not os.path.exists(ora.ORACLE_SO):
        ora.build_oracles()
    return ora, ora.OracleLib(), "port"


def cpu_reference_rate(cfg: dict, n_worlds: int, steps: int, warmup: int, threads: int, first_world: int = 0):
    """Step `n_worlds` sample worlds of the workload for `steps` steps on `threads` host threads.

    ctypes releases the GIL, so plain Python threads run the C step concurrently, one world each."""
    ora, lib, kind = _cpu_lib()
    st = scenes.plane_bodies(scenes.box_planes(cfg["L"]))
    pos, vel = scenes.sphere_worlds(n_worlds, cfg["n_spheres"], cfg["L"], first_world=first_world)
    worlds = [ora.World(lib, st, scenes.sphere_bodies(pos[k], vel[k])) for k in range(n_worlds)]
    for w in worlds:
        w.record_events(False)

    def run(nsteps):
        it = iter(worlds)
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    w = next(it, None)
                if w is None:
                    return
                w.step(scenes.DT_60HZ, nsteps, 0)
        ts = [threading.Thread(target=worker) for _ in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    if warmup:
        run(warmup)
    el = run(steps)
    for w in worlds:
        w.close()
    return n_worlds * cfg["n_spheres"] * steps / el, el, kind


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    # Size the sample so that the TIMED region lasts >= ~1.5 s whatever K is (a 40 ms region moved the ratio by
    # +-50 % between boxes): one world-step of 256 spheres costs ~1.3 ms on one core, so every thread gets
    # ceil(1.5 s / (K x 1.3 ms)) worlds, at least 2 and at most 64; a "step" is one physics_step of every
    # sample world.  K = 20 -> 58 worlds per thread, K = 1000 -> 2.
    steps = args.steps
    per_thread = int(min(64, max(2, -(-1.5 // (max(steps, 1) * 1.3e-3)))))
    n_sample = threads * per_thread
    if cfg.get("large"):
        # the reference sweep is O(N^2): ~1 minute per step at 65,536 bodies.  One world, one thread
        # (the reference is single-threaded), capped at 3 steps for c2; c4 is not runnable at all.
        if cfg["n_spheres"] > 100000:
            print(json.dumps({"impl": "reference", "unavailable": "reference physics_step is O(N^2): days per step at 4M bodies"}), flush=True)
            return
        threads, n_sample, steps = 1, 1, min(args.steps, 3)
        args.warmup = 0
    rate, el, kind = cpu_reference_rate(cfg, n_sample, steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": el / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": cfg["desc"], "n_spheres": cfg["n_spheres"], "dt": float(scenes.MAX_DELTA_TIME),
                   "sample": f"{n_sample} worlds x {cfg['n_spheres']} spheres, {threads} host threads"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n_sample} worlds x {steps} steps (of {cfg['n_worlds']} per GPU)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the CUDA path
# ------------------------------------------------------------------------------------------------

E2E_API_NOTE = ("pg_worlds_step_host_packed(dt, 1, state) [--e2e-layout split: pg_worlds_step_host(dt, 1, pos, vel, at_rest)]: pinned host "
                "state in and out every step; 4 world chunks (small first and last), two DMA queues per direction, the step kernel reads and writes the staging "
                "buffers itself, replayed as one CUDA graph")


def _set_affinity(local_rank: int):
    """Bind the rank to the CPUs NVML reports as closest to its GPU, so that its pinned host buffers are
    allocated (first touch) and copied NUMA-local; a no-op where NVML or the topology says nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def run_large_arm(args, cfg):
    """One large world per rank (replicas under N > 1: the slab decomposition is not built yet)."""
    import torch
    import torch.distributed as dist
    from crux_b200 import gpu

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, L = cfg["n_spheres"], cfg["L"]
    if cfg.get("level"):
        statics, p0, v0, players = scenes.level_world(n, L, seed_world=rank, with_player=True)
        world = gpu.GpuLargeWorld(n, len(statics), device=local_rank)
        world.set(gpu.prepare(statics), p0, v0)
        pl = gpu.prepare(players)
        pl["node_xform"][0] = gpu.node_transform(pl["position"][0], pl["rotation"][0], pl["scale"][0], pl["parent_xform"][0])
        world.set_players(pl)
    else:
        pos, vel = scenes.sphere_worlds(1, n, L, first_world=rank)
        statics = scenes.plane_bodies(scenes.box_planes(L))
        world = gpu.GpuLargeWorld(n, len(statics), device=local_rank)
        world.set(statics, pos[0], vel[0])
    dt = float(scenes.DT_60HZ)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    world.step(dt, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = world.launches
    ev_ms = 0.0
    kms = {}
    iters = []
    proven = True
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        world.step(dt, 1)
        ev_ms += world.last_step_ms()
        ex = world.exactness()
        iters.append(ex["iterations"])
        proven = proven and ex["proven_exact"]
    barrier()
    clocks = sampler.stop()
    gpu_launches = world.launches - launches0
    # per-kernel breakdown: a few more steps AFTER the timed region with the library's per-kernel event pairs
    # switched on (they are off by default: they double the API calls of a launch-bound pipeline)
    breakdown_steps = min(8, args.steps)
    world.set_timing(True)
    for _ in range(breakdown_steps):
        world.step(dt, 1)
        for k, v in world.kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v
    world.set_timing(False)
    kms = {k: v * args.steps / breakdown_steps for k, v in kms.items()}      # scaled to the timed region's step count
    t = torch.tensor([ev_ms], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ev_ms_max = float(t.item())
    value = n * world_size * args.steps / (ev_ms_max * 1e-3)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hp = gpu.PinnedArray((n, 3), np.float32); hv = gpu.PinnedArray((n, 3), np.float32); hr = gpu.PinnedArray((n,), np.uint8)
    world.download_state(hp.array, hv.array, hr.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        world.upload_state(hp.array, hv.array, hr.array)
        world.step(dt, 1)
        world.download_state(hp.array, hv.array, hr.array)
    barrier()
    e2e_value = n * world_size * e2e_steps / (time.perf_counter() - t0)
    stats = world.stats()
    hp.free(); hv.free(); hr.free()

    if rank == 0:
        peak, peak_src = read_peaks()
        per_kernel = {}
        traffic = {}
        rj = os.path.join(ROOT, "profiles", "roofline_large_r02.json")
        if os.path.exists(rj) and n == 4194304:             # the capture is of the 4 M-body world
            try:
                traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in json.load(open(rj))["kernels"].items()
                           if v.get("dram_bytes_read") is not None}
            except Exception:
                traffic = {}
        for k, b in LARGE_KERNEL_BYTES.items():
            ms = kms.get(k, 0.0) / args.steps
            if ms > 0:
                per_kernel[k] = {"ms": ms, "GBps": b * n / (ms * 1e-3) / 1e9, "frac": b * n / (ms * 1e-3) / 1e9 / peak, "traffic": traffic.get(k)}
        dom = max(kms, key=kms.get)
        dom_ms = kms[dom] / args.steps
        dom_bytes = LARGE_KERNEL_BYTES.get(dom, 96) * n      # k_eval / k_check: 96 B per contact is the survey figure; per body as an upper bound
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": cfg["desc"], "n_spheres": n, "n_static": int(len(statics)), "dt": float(scenes.MAX_DELTA_TIME),
                       "parallelism": "replicas only" if world_size > 1 else "single world, 1 GPU",
                       "l2": "256 MiB flush between steps, outside the CUDA-event pairs"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "streaming_kernels": per_kernel,
                         "note": "the dominant kernel is the order-preserving fixed point (latency-bound, few active pairs per iteration); "
                                 "the streaming kernels are listed with their own HBM fractions"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 25, "d2h_bytes_per_step": n * 25, "steps": e2e_steps,
                    "api": "pg_large_upload_state + pg_large_step(1) + pg_large_download_state, pinned host buffers"},
            "gpu_launches": int(gpu_launches),
            "fixed_point": {"iterations_mean": float(np.mean(iters)), "iterations_max": int(max(iters)), "proven_exact_every_step": bool(proven)},
            "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
            "kernel_ms_note": "per-kernel CUDA-event pairs from extra steps after the timed region (pg_large_set_timing)",
            "stats": stats,
        }
        print(json.dumps(line), flush=True)
    world.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist
    from crux_b200 import gpu, shard

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _set_affinity(local_rank)
    comm = None
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # the path's one collective lives in the library: its own NCCL communicator, bootstrapped with a 128-byte
        # id that torch.distributed carries from rank 0 to the others (plumbing only)
        comm = gpu.Comm.from_torch_distributed(local_rank)

    nw, n, L = cfg["n_worlds"], cfg["n_spheres"], cfg["L"]
    statics = scenes.plane_bodies(scenes.box_planes(L))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    spl = max(1, min(args.steps_per_launch, args.steps))
    dt = float(scenes.DT_60HZ)

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_worlds(first_world, count):
        pos, vel = scenes.sphere_worlds(count, n, L, first_world=first_world)
        w = gpu.GpuWorlds(count, len(statics), n, 0, max_events_per_world=0, device=local_rank)
        w.set_spheres(0, statics, pos, vel)
        return w

    def launch(worlds, k):
        """One timed unit: L2 flush (outside the clock), then k steps in one launch + the statistics all-reduce,
        both on the library's stream inside ONE CUDA-event pair (pg_worlds_last_step_ms)."""
        flush.zero_()
        torch.cuda.synchronize()
        worlds.step(dt, k)
        worlds.stats_allreduce(comm)                                        # NCCL inside the library, no host sync
        worlds.sync()
        return worlds.last_step_ms()

    def timed(worlds, steps, warmup):
        done = 0
        while done < warmup:
            k = min(spl, warmup - done)
            launch(worlds, k)
            done += k
        barrier()
        ev_ms, per = 0.0, []
        t0 = time.perf_counter()
        done = 0
        while done < steps:
            k = min(spl, steps - done)
            ms = launch(worlds, k)
            ev_ms += ms
            per.append(ms / k)
            done += k
        barrier()
        return max_over_ranks(ev_ms), per, time.perf_counter() - t0

    # ---- weak scaling (the driver's line): every rank owns nw worlds
    bodies = nw * n
    worlds = make_worlds(rank * nw, nw)
    sampler = ClockSampler(local_rank)
    launches0 = worlds.launches
    sampler.start()
    ev_ms_max, launch_ms, wall = timed(worlds, args.steps, args.warmup)
    gpu_launches = worlds.launches - launches0
    stats_sum = worlds.stats_allreduce(comm, want_host=True)                # summed over all ranks
    value = bodies * world_size * args.steps / (ev_ms_max * 1e-3)

    # ---- the hard regime: the same batch after its worlds have thermalised (pre-advanced, untimed)
    therm = None
    if args.thermalise_steps > 0:
        worlds.step(dt, args.thermalise_steps)
        worlds.sync()
        ev2, per2, _ = timed(worlds, args.steps, 0)
        therm = {"value": bodies * world_size * args.steps / (ev2 * 1e-3), "ms_per_step": ev2 / args.steps,
                 "pre_advanced_steps": args.thermalise_steps + args.warmup + args.steps}
    clocks = sampler.stop()
    worlds.close()

    # ---- strong scaling: BASELINE's named batch (4096 worlds in total) split over the ranks
    strong = None
    if world_size > 1:
        first, count = shard.shard_range(cfg["n_worlds"], rank, world_size)
        sw = make_worlds(first, count)
        ev3, _, _ = timed(sw, args.steps, args.warmup)
        strong = {"value": cfg["n_worlds"] * n * args.steps / (ev3 * 1e-3), "ms_per_step": ev3 / args.steps,
                  "worlds_total": cfg["n_worlds"], "worlds_per_gpu": count, "scaling": "strong"}
        sw.close()

    # ---- end to end through the host-buffer API: H2D state, one step, D2H state, every step
    worlds = make_worlds(rank * nw, nw)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hp = gpu.PinnedArray((nw, n, 3), np.float32); hv = gpu.PinnedArray((nw, n, 3), np.float32)
    hr = gpu.PinnedArray((nw, n), np.uint8)
    worlds.download_state(0, nw, hp.array, hv.array, hr.array)
    packed = args.e2e_layout == "packed"
    if packed:
        # one pinned buffer, one record per world (pos | vel | at_rest): a chunk of worlds is one DMA each way
        rec = worlds.packed_dtype()
        hs = gpu.PinnedArray((nw,), rec)
        hs.array["pos"][:] = hp.array; hs.array["vel"][:] = hv.array; hs.array["at_rest"][:] = hr.array

        def host_step():
            worlds.step_host_packed(dt, 1, hs.array)
    else:
        def host_step():
            worlds.step_host(dt, 1, hp.array, hv.array, hr.array)
    for _ in range(3):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()                                                      # H2D + step + D2H, synchronous
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_value = bodies * world_size * e2e_steps / e2e_wall
    bytes_per_body = (rec.itemsize / n) if packed else 25
    h2d = int(bodies * bytes_per_body)
    d2h = int(bodies * bytes_per_body)
    checksum = float(np.abs(hs.array["pos"]).sum()) if packed else float(np.abs(hp.array).sum())
    if packed:
        hs.free()
    hp.free(); hv.free(); hr.free()

    if rank == 0:
        peak, peak_src = read_peaks()
        avg_launch_ms = ev_ms_max / max(1, len(launch_ms))
        algo_bytes = ALGO_BYTES_PER_BODY_STEP * bodies * spl
        achieved = algo_bytes / (avg_launch_ms * 1e-3) / 1e9
        traffic = None
        issue = {}
        for name in ("roofline_r02.json", "roofline_r01.json"):
            rj = os.path.join(ROOT, "profiles", name)
            if not os.path.exists(rj):
                continue
            try:
                prof = json.load(open(rj))
                traffic = prof.get("dram_bytes_per_step")
                traffic = traffic * spl if traffic else None
                issue = {k: prof[k] for k in ("sm_issue_utilisation", "warp_instructions_per_world_step", "icache_hit_rate",
                                              "gcc_instruction_request_rate_of_peak") if k in prof}
                issue["source"] = "profiles/" + name + " (ncu capture, not measured live)"
                break
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": cfg["desc"], "worlds_per_gpu": nw, "n_spheres": n, "n_planes": 6,
                       "dt": float(scenes.MAX_DELTA_TIME), "steps_per_launch": spl, "parallelism": f"worlds-sharded x{world_size}",
                       "collective": "8-double statistics all-reduce (NCCL, inside libphysics_gpu.so) once per launch, inside the timed CUDA-event pair",
                       "l2": "256 MiB flush between launches, outside the CUDA-event pairs"},
            "roofline": {"bound": "sm_issue", "kernel": "k_step_spheres<8,16>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "issue": issue,
                         "note": "achieved/peak/frac are the LIVE HBM figures (algorithmic 52 B per body-step over the measured launch time); "
                                 "they are a secondary reading: state lives in shared memory for the whole launch and the kernel is bound by SM "
                                 "issue slots and instruction fetch (SURVEY.md section 8d), see profiles/k_step_spheres_r02.md"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": E2E_API_NOTE},
            "gpu_launches": int(gpu_launches),
            "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
            "launch_ms_per_step_min_max": [min(launch_ms), max(launch_ms)],
            "stats": {k: stats_sum[k] for k in ("bodies", "kinetic", "contacts", "at_rest", "nonfinite")},
            "state_checksum": checksum,
        }
        if therm:
            line["value_thermalised"] = therm["value"]
            line["thermalised"] = therm
        if strong:
            line["strong"] = strong
        if world_size == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            threads = min(cores, 64)
            # ~10 s of CPU work per thread: 6 worlds per thread x 1000 steps at ~1.3 ms per world-step
            cpu_steps = 1000
            rate, el, kind = cpu_reference_rate(cfg, threads * 6, cpu_steps, 0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{threads * 6} worlds x {cpu_steps} steps of the same workload, {el:.1f} s wall"}
        print(json.dumps(line), flush=True)

    worlds.close()
    if comm:
        comm.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 1000 for c3, 24 for the large-world workloads)")
    ap.add_argument("--warmup", type=int, default=None, help="untimed steps first (default: 20 for c3, 4 for the large-world workloads)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--steps-per-launch", type=int, default=10)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-layout", default="packed", choices=["packed", "split"],
                    help="host layout of the end-to-end leg: one packed record per world, or three split arrays")
    ap.add_argument("--thermalise-steps", type=int, default=1000,
                    help="c3: after the timed region, pre-advance the batch this many steps (untimed) and time K steps again "
                         "(value_thermalised); 0 to skip")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = 24 if cfg.get("large") else 1000
    if args.warmup is None:
        args.warmup = 4 if cfg.get("large") else 20
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    elif cfg.get("large"):
        run_large_arm(args, cfg)
    else:
        run_gpu_arm(args, cfg)


if __name__ == "__main__":
    main()
