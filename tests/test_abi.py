"""The C-ABI library: it loads, exports every symbol include/physics_gpu.h declares, lays records out
as the header says, and refuses to work without a GPU (no compute is attempted here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from crux_b200 import gpu, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "physics_gpu.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 40
    dll = C.CDLL(gpu.GPU_SO)
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, f"declared in physics_gpu.h but not exported: {missing}"
    bound = {n for n, _, _ in gpu.API}
    assert set(names) == bound, f"python binding table and header disagree: {set(names) ^ bound}"


def test_record_layouts_match_header():
    assert scenes.BODY_DTYPE.itemsize == 304 and scenes.CORE_DTYPE.itemsize == 232
    assert gpu.PG_EVENT_DTYPE.itemsize == 32
    offs = {n: scenes.BODY_DTYPE.fields[n][1] for n in scenes.BODY_DTYPE.names}
    assert (offs["position"], offs["rotation"], offs["scale"], offs["velocity"], offs["restitution"], offs["shape"]) == (0, 12, 24, 36, 48, 52)
    assert (offs["collider_type"], offs["node_xform"], offs["parent_xform"], offs["euler_raw"], offs["euler_node"]) == (80, 104, 168, 232, 268)
    # PgDebugDraw: kind, cls, index, n_indices, 24 floats
    d = scenes.DEBUG_DTYPE
    assert d.itemsize == 112 and [d.fields[n][1] for n in ("kind", "cls", "index", "n_indices", "data")] == [0, 4, 8, 12, 16]
    header = open(HEADER).read()
    for name, value in (("PG_DEBUG_NONE", scenes.DEBUG_NONE), ("PG_DEBUG_AABB_LINES", scenes.DEBUG_AABB_LINES), ("PG_DEBUG_SPHERE_MODEL", scenes.DEBUG_SPHERE_MODEL),
                        ("PG_DEBUG_CAPSULE_MODEL", scenes.DEBUG_CAPSULE_MODEL), ("PG_DEBUG_CAPSULE_HOST", scenes.DEBUG_CAPSULE_HOST)):
        assert re.search(rf"\b{name} = {value}\b", header), name


def test_host_helpers_agree_with_oracle(oracle_lib):
    """pg_body_prepare / pg_node_transform are host code (libm): check them against the port."""
    import ctypes
    rng = np.random.default_rng(4)
    b = scenes.new_bodies(16)
    b["rotation"] = rng.choice([0, 90, 37.5, -120, 180], (16, 3)).astype(np.float32)
    gpu.prepare(b)
    for k in range(16):
        raw = np.zeros(9, np.float32)
        oracle_lib.dll.ora_euler_xyz(b["rotation"][k].ctypes.data_as(ctypes.c_void_p), raw.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(raw.view(np.uint32), b["euler_raw"][k].view(np.uint32))
        pos = rng.uniform(-3, 3, 3).astype(np.float32); sc = rng.uniform(0.5, 2, 3).astype(np.float32)
        par = gpu.node_transform(rng.uniform(-1, 1, 3), [0, 90, 0], [1, 1, 1])
        got = gpu.node_transform(pos, b["rotation"][k], sc, par)
        want = np.zeros(16, np.float32)
        oracle_lib.dll.ora_node_transform(pos.ctypes.data_as(ctypes.c_void_p), b["rotation"][k].ctypes.data_as(ctypes.c_void_p),
                                          sc.ctypes.data_as(ctypes.c_void_p), par.ctypes.data_as(ctypes.c_void_p),
                                          want.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_no_gpu_means_loud_failure():
    L = gpu.lib()
    if L.pg_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(gpu.PgError):
        gpu.GpuWorlds(1, 1, 1)
    assert b"cuda" in L.pg_last_error().lower()
