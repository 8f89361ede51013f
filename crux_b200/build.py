"""Build the native libraries in-tree (sm_100a only).

  crux_b200/lib/libphysics_gpu.so   CUDA kernels + the C-ABI of include/physics_gpu.h
  crux_b200/lib/libcrux_physics.so  the reference-named C host API (crux_b200/csrc/host/*.c), linked
                                    against libphysics_gpu.so

`-fmad=false` is load-bearing: the parity contract is bit-exactness against a non-FMA x86 build
(see crux_b200/csrc/pg_math.cuh).  nvcc cross-compiles without a GPU present.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
GPU_SO = os.path.join(LIB, "libphysics_gpu.so")
HOST_SO = os.path.join(LIB, "libcrux_physics.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
    "-Xptxas", "-v",
]
EXTRA_INCLUDES: list[str] = []
CU_SOURCES = ["physics_gpu.cu", "pg_general.cu", "pg_spheres.cu", "pg_large.cu", "pg_comm.cu", "pg_debug.cu"]
HOST_SOURCES = ["host/crux_world.c", "host/crux_colliders.c", "host/crux_scene.c"]
HEADLESS = os.path.join(LIB, "crux_headless")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers() -> list[str]:
    hs = [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hostdir = os.path.join(CSRC, "host")
    if os.path.isdir(hostdir):
        hs += [os.path.join(hostdir, f) for f in os.listdir(hostdir) if f.endswith(".h")]
    return hs


def build_gpu(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force or _newer(GPU_SO, srcs + _headers() + [os.path.abspath(__file__)]):
        objs = []
        logs = []
        for s in srcs:
            o = os.path.join(LIB, os.path.basename(s).replace(".cu", ".o"))
            tu_log = o[:-2] + ".ptxas.log"         # per translation unit, so that ptxas.log always covers all of them
            if force or _newer(o, [s] + _headers() + [os.path.abspath(__file__)]) or not os.path.exists(tu_log):
                r = subprocess.run([NVCC, *NVCC_FLAGS, *EXTRA_INCLUDES, "-c", s, "-o", o], capture_output=True, text=True)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed on {s}")
                with open(tu_log, "w") as f:
                    f.write(f"==== {os.path.basename(s)} ====\n" + r.stderr)
            logs.append(open(tu_log).read())
            objs.append(o)
        r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", GPU_SO, *objs, "-ldl"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc link failed")
        with open(os.path.join(LIB, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
        if verbose:
            sys.stderr.write("\n".join(logs))
    return GPU_SO


def build_host(force: bool = False) -> str | None:
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not srcs:
        return None
    if force or _newer(HOST_SO, srcs + _headers() + [GPU_SO]):
        cmd = ["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra", "-fPIC", "-shared",
               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CSRC, "host"),
               "-I" + os.path.join(ROOT, "third_party"),
               "-o", HOST_SO, *srcs, "-L" + LIB, "-lphysics_gpu", "-Wl,-rpath,$ORIGIN", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed on the host layer")
    main_c = os.path.join(CSRC, "host", "crux_headless.c")
    if os.path.exists(main_c) and (force or _newer(HEADLESS, [main_c, HOST_SO] + _headers())):
        cmd = ["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-rdynamic", "-DCRUX_HAVE_UPLOAD_STATS",
               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CSRC, "host"),
               "-o", HEADLESS, main_c, "-L" + LIB, "-lcrux_physics", "-lphysics_gpu", "-Wl,-rpath,$ORIGIN", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed on crux_headless")
    return HOST_SO


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_gpu(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print(GPU_SO)
