"""ctypes bindings for the two CPU checkers under oracle/ (test infrastructure, never product).

`RefLib`    -> oracle/_ref/libcrux_ref.so      the unmodified reference sources + ref_harness.c
`OracleLib` -> oracle/_build/libcrux_oracle.so the plain-C restatement (crux_oracle.c)

Both export the same flat API (prefix `ref_` / `ora_`), so a `World` works against either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from crux_b200.scenes import CORE_DTYPE, EVENT_DTYPE, BODY_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libcrux_ref.so")
REF_O0_SO = os.path.join(ORACLE_DIR, "_ref", "libcrux_ref_O0.so")
REF_DEBUG_SO = os.path.join(ORACLE_DIR, "_ref", "libcrux_ref_debug.so")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libcrux_oracle.so")
REFERENCE_ROOT = "/root/reference"


def build_oracles() -> None:
    """Compile the checkers (restatement always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def to_core(bodies: np.ndarray) -> np.ndarray:
    """BODY_DTYPE (PgBody) or CORE_DTYPE records -> contiguous CORE_DTYPE (OraBody) records."""
    bodies = np.atleast_1d(bodies)
    out = np.zeros(bodies.shape, dtype=CORE_DTYPE)
    for name in CORE_DTYPE.names:
        out[name] = bodies[name]
    return out


def from_core(core: np.ndarray, like: np.ndarray | None = None) -> np.ndarray:
    out = np.zeros(core.shape, dtype=BODY_DTYPE) if like is None else like.copy()
    for name in CORE_DTYPE.names:
        out[name] = core[name]
    return out


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _Lib:
    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path)
        f = self._f
        f("world_create", C.c_void_p, [])
        f("world_destroy", None, [C.c_void_p])
        f("world_add_many", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int])
        f("world_remove", C.c_int, [C.c_void_p, C.c_int, C.c_int])
        f("world_counts", None, [C.c_void_p, C.c_void_p])
        f("world_read", None, [C.c_void_p, C.c_int, C.c_void_p])
        f("world_poke", None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p])
        f("world_refresh_nodes", None, [C.c_void_p])
        f("world_step", None, [C.c_void_p, C.c_float, C.c_int, C.c_int])
        f("world_record_events", None, [C.c_void_p, C.c_int])
        f("events_count", C.c_int, [C.c_void_p])
        f("events_read", None, [C.c_void_p, C.c_void_p])
        f("events_clear", None, [C.c_void_p])
        f("max_move", C.c_float, [C.c_void_p, C.c_float, C.c_float])
        f("min_dist", C.c_float, [C.c_void_p, C.c_void_p, C.c_float])
        f("interval_collision", C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p])
        f("narrow_phase", C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p])
        f("resolve", C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float])
        f("visit", C.c_int, [C.c_void_p, C.c_void_p, C.c_float])

    def _f(self, name, restype, argtypes):
        fn = getattr(self.dll, f"{self.prefix}_{name}")
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, name, fn)

    # ---- per-pair known-answer helpers (records are 1-element CORE arrays) ----
    def k_min_dist(self, a, b, t):
        a, b = to_core(a), to_core(b)
        return np.float32(self.min_dist(_ptr(a), _ptr(b), np.float32(t)))

    def k_max_move(self, a, t0, t1):
        a = to_core(a)
        return np.float32(self.max_move(_ptr(a), np.float32(t0), np.float32(t1)))

    def k_interval(self, a, b, t0, t1):
        a, b = to_core(a), to_core(b)
        hit = np.zeros(1, dtype=np.float32)
        r = self.interval_collision(_ptr(a), _ptr(b), np.float32(t0), np.float32(t1), _ptr(hit))
        return int(r), np.float32(hit[0])

    def k_narrow(self, a, b, dt):
        a, b = to_core(a), to_core(b)
        hit = np.zeros(1, dtype=np.float32)
        r = self.narrow_phase(_ptr(a), _ptr(b), np.float32(dt), _ptr(hit))
        return int(r), np.float32(hit[0])

    def k_resolve(self, a, b, hit_time, dt):
        a, b = to_core(a), to_core(b)
        r = self.resolve(_ptr(a), _ptr(b), np.float32(hit_time), np.float32(dt))
        return int(r), a, b

    def k_visit(self, a, b, dt):
        a, b = to_core(a), to_core(b)
        r = self.visit(_ptr(a), _ptr(b), np.float32(dt))
        return int(r), a, b


class RefLib(_Lib):
    def __init__(self, path: str = REF_SO):
        super().__init__(path, "ref")


class RefDebugLib(_Lib):
    """The reference physics + its UNMODIFIED debug renderer behind a recording OpenGL stub
    (oracle/ref_debug_harness.c).  `debug_init` / `debug_render` return the log of one
    physics_debug_renderer_init / physics_debug_render call as (what, count, tag, data) tuples."""
    VERTEX_DATA, INDEX_DATA, VERTEX_SUBDATA, MODEL_MAT4, DRAW = 1, 2, 3, 4, 5

    def __init__(self, path: str = REF_DEBUG_SO):
        super().__init__(path, "ref")
        d = self.dll
        d.ref_debug_init.restype = C.c_int; d.ref_debug_init.argtypes = [C.c_void_p]
        d.ref_debug_render.restype = C.c_int; d.ref_debug_render.argtypes = [C.c_void_p]
        d.ref_debug_capture.restype = C.c_int
        d.ref_debug_capture.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]

    def _log(self, n):
        out = []
        buf = np.zeros(8192, np.uint32)
        what, count, tag = C.c_int(0), C.c_int(0), C.c_int(0)
        for k in range(n):
            m = self.dll.ref_debug_capture(k, C.byref(what), C.byref(count), C.byref(tag), _ptr(buf), len(buf))
            assert m >= 0, m
            words = buf[:m].copy()
            data = words if what.value == self.INDEX_DATA else words.view(np.float32)
            out.append((what.value, count.value, tag.value, data))
        return out

    def debug_init(self, world: "World"):
        return self._log(self.dll.ref_debug_init(world.h))

    def debug_render(self, world: "World"):
        return self._log(self.dll.ref_debug_render(world.h))


class OracleLib(_Lib):
    def __init__(self, path: str = ORACLE_SO):
        super().__init__(path, "ora")
        d = self.dll
        d.ora_set_threads.restype = None
        d.ora_set_threads.argtypes = [C.c_int]
        d.ora_world_step_grid.restype = None
        d.ora_world_step_grid.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        d.ora_world_pairs.restype = C.c_longlong
        d.ora_world_pairs.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_longlong]


class World:
    """One physics world living inside a checker library."""

    def __init__(self, lib: _Lib, statics=None, dynamics=None, players=None):
        self.lib = lib
        self.h = lib.world_create()
        for recs, as_player in ((statics, 0), (dynamics, 0), (players, 1)):
            if recs is not None and len(recs):
                self.add(recs, as_player)

    def add(self, recs, as_player=0):
        core = to_core(recs)
        return self.lib.world_add_many(self.h, _ptr(core), len(core), int(as_player))

    def remove(self, cls, index):
        self.lib.world_remove(self.h, cls, index)

    def counts(self):
        c = np.zeros(3, dtype=np.int32)
        self.lib.world_counts(self.h, _ptr(c))
        return tuple(int(x) for x in c)

    def read(self, cls) -> np.ndarray:
        out = np.zeros(self.counts()[cls], dtype=CORE_DTYPE)
        if len(out):
            self.lib.world_read(self.h, cls, _ptr(out))
        return out

    def poke(self, cls, index, rec):
        core = to_core(rec)
        self.lib.world_poke(self.h, cls, index, _ptr(core))

    def step(self, dt, nsteps=1, refresh_nodes=1):
        self.lib.world_step(self.h, np.float32(dt), int(nsteps), int(refresh_nodes))

    def record_events(self, on: bool):
        self.lib.world_record_events(self.h, int(on))

    def events(self, clear=True) -> np.ndarray:
        n = self.lib.events_count(self.h)
        out = np.zeros(n, dtype=EVENT_DTYPE)
        if n:
            self.lib.events_read(self.h, _ptr(out))
        if clear:
            self.lib.events_clear(self.h)
        return out

    def close(self):
        if self.h:
            self.lib.world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
