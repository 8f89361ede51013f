"""ctypes binding of libphysics_gpu.so (the C-ABI declared in include/physics_gpu.h).

Nothing here computes physics: every method forwards to the CUDA library.  Loading fails loudly if
the library has not been built (python -m crux_b200.build) -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scenes import BODY_DTYPE, CLASS_DYNAMIC, CLASS_PLAYER, CLASS_STATIC

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_SO = os.environ.get("CRUX_PG_LIB") or os.path.join(HERE, "lib", "libphysics_gpu.so")

PG_EVENT_DTYPE = np.dtype([("world", "<i4"), ("step", "<i4"), ("event_type", "<i4"), ("a_class", "<i4"),
                           ("a_index", "<i4"), ("b_class", "<i4"), ("b_index", "<i4"), ("pass_", "<i4")])
PG_STATS_FIELDS = ("bodies", "kinetic", "px", "py", "pz", "contacts", "at_rest", "nonfinite")

# every symbol include/physics_gpu.h declares: (name, restype, argtypes)
_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
API = [
    ("pg_last_error", C.c_char_p, []),
    ("pg_device_count", _i, []),
    ("pg_body_prepare", None, [_vp, _i]),
    ("pg_node_transform", None, [_vp, _vp, _vp, _vp, _vp]),
    ("pg_host_alloc", _vp, [C.c_size_t]),
    ("pg_host_free", None, [_vp]),
    ("pg_worlds_create", _vp, [_i, _i, _i, _i, _i, _i]),
    ("pg_worlds_destroy", None, [_vp]),
    ("pg_worlds_count", _i, [_vp]),
    ("pg_worlds_set_bodies", _i, [_vp, _i, _i, _vp, _i]),
    ("pg_worlds_update_bodies", _i, [_vp, _i, _i, _i, _vp, _i]),
    ("pg_worlds_add_body", _i, [_vp, _i, _vp, _i]),
    ("pg_worlds_remove_body", _i, [_vp, _i, _i, _i]),
    ("pg_worlds_body_count", _i, [_vp, _i, _i]),
    ("pg_worlds_get_bodies", _i, [_vp, _i, _i, _vp, _i]),
    ("pg_worlds_set_spheres", _i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _f, _f]),
    ("pg_worlds_upload_state", _i, [_vp, _i, _i, _vp, _vp, _vp]),
    ("pg_worlds_download_state", _i, [_vp, _i, _i, _vp, _vp, _vp]),
    ("pg_worlds_step", _i, [_vp, _f, _i, _i]),
    ("pg_worlds_sync", _i, [_vp]),
    ("pg_worlds_step_host", _i, [_vp, _f, _i, _vp, _vp, _vp]),
    ("pg_worlds_packed_world_bytes", C.c_size_t, [_vp]),
    ("pg_worlds_step_host_packed", _i, [_vp, _f, _i, _vp]),
    ("pg_worlds_frame", _i, [_vp, _i, _vp, _vp, _i, _vp, _f, _i, _i, _vp, _vp, _vp, _vp, _ll, _vp, _vp]),
    ("pg_worlds_events", _ll, [_vp, _vp, _ll, _vp]),
    ("pg_worlds_stats", _i, [_vp, _vp, _vp]),
    ("pg_comm_unique_id", _i, [_vp]),
    ("pg_comm_init", _vp, [_vp, _i, _i, _i]),
    ("pg_comm_destroy", None, [_vp]),
    ("pg_comm_rank", _i, [_vp]),
    ("pg_comm_size", _i, [_vp]),
    ("pg_comm_nccl_version", _i, []),
    ("pg_worlds_stats_allreduce", _i, [_vp, _vp, _vp, _vp]),
    ("pg_worlds_launches", _ll, [_vp]),
    ("pg_worlds_last_step_ms", _f, [_vp]),
    ("pg_worlds_stream", _vp, [_vp]),
    ("pg_worlds_debug_geometry", _i, [_vp, _i, _vp, _i, _i]),
    ("pg_debug_sphere_mesh", None, [_f, _vp, _vp]),
    ("pg_debug_capsule_mesh", None, [_vp, _vp, _f, _vp, _vp]),
    ("pg_debug_box_mesh", None, [_vp, _vp, _vp, _vp]),
    ("pg_debug_capsule_model", None, [_vp, _vp]),
    ("pg_large_create", _vp, [_i, _i, _i]),
    ("pg_large_destroy", None, [_vp]),
    ("pg_large_set", _i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _f, _f]),
    ("pg_large_set_players", _i, [_vp, _vp, _i]),
    ("pg_large_get_players", _i, [_vp, _vp, _i]),
    ("pg_large_upload_state", _i, [_vp, _vp, _vp, _vp]),
    ("pg_large_download_state", _i, [_vp, _vp, _vp, _vp]),
    ("pg_large_step", _i, [_vp, _f, _i]),
    ("pg_large_sync", _i, [_vp]),
    ("pg_large_pairs", _ll, [_vp, _f, _vp, _ll]),
    ("pg_large_events", _ll, [_vp, _vp, _ll, _vp]),
    ("pg_large_stats", _i, [_vp, _vp, _vp]),
    ("pg_large_launches", _ll, [_vp]),
    ("pg_large_last_step_ms", _f, [_vp]),
    ("pg_large_kernel_count", _i, []),
    ("pg_large_kernel_name", C.c_char_p, [_i]),
    ("pg_large_kernel_ms", _f, [_vp, _i]),
    ("pg_large_set_timing", _i, [_vp, _i]),
    ("pg_large_exactness", _i, [_vp, _vp, _vp, _vp, _vp]),
    ("pg_pair_max_move", _i, [_vp, _i, _f, _f, _vp]),
    ("pg_pair_min_dist", _i, [_vp, _vp, _i, _vp, _vp]),
    ("pg_pair_interval_collision", _i, [_vp, _vp, _i, _f, _f, _vp, _vp]),
    ("pg_pair_narrow_phase", _i, [_vp, _vp, _i, _f, _vp, _vp]),
    ("pg_pair_resolve", _i, [_vp, _vp, _i, _vp, _f]),
    ("pg_pair_visit", _i, [_vp, _vp, _i, _f, _vp]),
]


class PgError(RuntimeError):
    pass


class PgUnproven(PgError):
    """pg_large_step returned PG_ERR_UNPROVEN: the steps ran, but equality with the reference's sequential
    sweep is not established for at least one of them."""


PG_ERR_UNPROVEN = -5


_lib = None


def lib() -> C.CDLL:
    """Load libphysics_gpu.so and bind every declared symbol (raises if the build is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(GPU_SO):
            raise PgError(f"{GPU_SO} not built: run `python -m crux_b200.build` (needs nvcc). "
                          "There is no CPU fallback.")
        dll = C.CDLL(GPU_SO)
        for name, restype, argtypes in API:
            fn = getattr(dll, name)          # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = dll
    return _lib


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, what: str):
    if rc is None or (isinstance(rc, int) and rc < 0):
        msg = lib().pg_last_error()
        cls = PgUnproven if rc == PG_ERR_UNPROVEN else PgError
        raise cls(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
    return rc


def _state_array(a, shape, dtype, what: str, writable: bool = False, optional: bool = False):
    """The library copies exactly prod(shape) items of `dtype` through the raw pointer: refuse anything else
    (a float64, sliced, short or read-only array would corrupt the heap silently)."""
    if a is None:
        if optional:
            return None
        raise PgError(f"{what}: array required")
    if not isinstance(a, np.ndarray):
        raise PgError(f"{what}: numpy array expected, got {type(a).__name__}")
    if a.dtype != np.dtype(dtype):
        raise PgError(f"{what}: dtype {a.dtype} (need {np.dtype(dtype)})")
    if tuple(a.shape) != tuple(shape):
        raise PgError(f"{what}: shape {tuple(a.shape)} (need {tuple(shape)})")
    if not a.flags["C_CONTIGUOUS"]:
        raise PgError(f"{what}: array must be C-contiguous")
    if writable and not a.flags["WRITEABLE"]:
        raise PgError(f"{what}: array must be writable")
    return a


def prepare(bodies: np.ndarray) -> np.ndarray:
    """Fill the two host-computed Euler matrices of BODY_DTYPE records (in place) and return them."""
    bodies = np.ascontiguousarray(bodies, dtype=BODY_DTYPE)
    lib().pg_body_prepare(_ptr(bodies), len(bodies))
    return bodies


def node_transform(pos, rot_deg, scale, parent=None) -> np.ndarray:
    out = np.zeros(16, np.float32)
    p = np.ascontiguousarray(pos, np.float32); r = np.ascontiguousarray(rot_deg, np.float32)
    s = np.ascontiguousarray(scale, np.float32)
    par = None if parent is None else np.ascontiguousarray(parent, np.float32)
    lib().pg_node_transform(_ptr(p), _ptr(r), _ptr(s), _ptr(par), _ptr(out))
    return out


def debug_sphere_mesh(radius: float):
    """Vertices (441, 3) and triangle indices (2280,) of physics_debug_sphere_init (debug_renderer.c:371-432)."""
    v = np.zeros((441, 3), np.float32); idx = np.zeros(2280, np.uint32)
    lib().pg_debug_sphere_mesh(np.float32(radius), _ptr(v), _ptr(idx))
    return v, idx


def debug_capsule_mesh(seg_a, seg_b, radius: float):
    """Vertices (882, 3) and triangle indices (4920,) of physics_debug_capsule_init (debug_renderer.c:469-616)."""
    a = np.ascontiguousarray(seg_a, np.float32); b = np.ascontiguousarray(seg_b, np.float32)
    v = np.zeros((882, 3), np.float32); idx = np.zeros(4920, np.uint32)
    lib().pg_debug_capsule_mesh(_ptr(a), _ptr(b), np.float32(radius), _ptr(v), _ptr(idx))
    return v, idx


def debug_box_mesh(center, extents):
    """Corner vertices (8, 3) and line indices (24,) of physics_debug_AABB_init (debug_renderer.c:311-337)."""
    c = np.ascontiguousarray(center, np.float32); e = np.ascontiguousarray(extents, np.float32)
    v = np.zeros((8, 3), np.float32); idx = np.zeros(24, np.uint32)
    lib().pg_debug_box_mesh(_ptr(c), _ptr(e), _ptr(v), _ptr(idx))
    return v, idx


def debug_capsule_model(body: np.ndarray) -> np.ndarray:
    """Model matrix (16 floats, column-major) of a capsule drawn without a scene node (debug_renderer.c:53-59)."""
    rec = np.ascontiguousarray(np.atleast_1d(body)[:1])
    assert rec.dtype == BODY_DTYPE
    m = np.zeros(16, np.float32)
    lib().pg_debug_capsule_model(_ptr(rec), _ptr(m))
    return m


class PinnedArray:
    """A numpy view over page-locked host memory from pg_host_alloc."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().pg_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise PgError("pg_host_alloc failed: " + lib().pg_last_error().decode())
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().pg_host_free(self.ptr)
            self.ptr = None


class Comm:
    """An NCCL communicator owned by the library (PgComm): one per process/GPU."""

    def __init__(self, unique_id: bytes, rank: int, n_ranks: int, device: int):
        if len(unique_id) != 128:
            raise PgError("Comm: the NCCL unique id is 128 bytes")
        self.L = lib()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.h = self.L.pg_comm_init(buf, rank, n_ranks, device)
        if not self.h:
            raise PgError("pg_comm_init failed: " + self.L.pg_last_error().decode())
        self.rank, self.n_ranks = rank, n_ranks

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().pg_comm_unique_id(buf), "pg_comm_unique_id")
        return buf.raw

    @staticmethod
    def from_torch_distributed(device: int) -> "Comm":
        """Rank 0 makes the id, torch.distributed (any backend) carries the 128 bytes to the others."""
        import torch
        import torch.distributed as dist
        rank, n = dist.get_rank(), dist.get_world_size()
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return Comm(box[0], rank, n, device)

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_comm_destroy(self.h)
            self.h = None


class GpuWorlds:
    """A batch of physics worlds resident on one GPU (PgWorlds handle)."""

    def __init__(self, n_worlds: int, max_static: int, max_dynamic: int, max_player: int = 0,
                 max_events_per_world: int = 0, device: int = 0):
        self.L = lib()
        self.n_worlds = n_worlds
        self.max_dynamic = max_dynamic
        self.h = self.L.pg_worlds_create(n_worlds, max_static, max_dynamic, max_player,
                                         max_events_per_world, device)
        if not self.h:
            raise PgError("pg_worlds_create failed: " + self.L.pg_last_error().decode())

    # -- general records
    def set_bodies(self, world: int, cls: int, bodies: np.ndarray):
        b = prepare(bodies)
        _check(self.L.pg_worlds_set_bodies(self.h, world, cls, _ptr(b), len(b)), "pg_worlds_set_bodies")

    def update_bodies(self, world: int, cls: int, first: int, bodies: np.ndarray):
        """Dirty-range upload: overwrite records [first, first+len(bodies)) of a class, verbatim."""
        b = np.ascontiguousarray(bodies, dtype=BODY_DTYPE)
        _check(self.L.pg_worlds_update_bodies(self.h, world, cls, first, _ptr(b), len(b)), "pg_worlds_update_bodies")

    def add_body(self, world: int, body: np.ndarray, as_player: bool = False) -> int:
        b = prepare(np.atleast_1d(body))
        return _check(self.L.pg_worlds_add_body(self.h, world, _ptr(b), int(as_player)), "pg_worlds_add_body")

    def remove_body(self, world: int, cls: int, index: int):
        _check(self.L.pg_worlds_remove_body(self.h, world, cls, index), "pg_worlds_remove_body")

    def body_count(self, world: int, cls: int) -> int:
        return _check(self.L.pg_worlds_body_count(self.h, world, cls), "pg_worlds_body_count")

    def get_bodies(self, world: int, cls: int) -> np.ndarray:
        n = self.body_count(world, cls)
        out = np.zeros(max(n, 1), dtype=BODY_DTYPE)
        got = _check(self.L.pg_worlds_get_bodies(self.h, world, cls, _ptr(out), len(out)), "pg_worlds_get_bodies")
        return out[:got]

    # -- sphere batches
    def set_spheres(self, first: int, statics: np.ndarray, pos: np.ndarray, vel: np.ndarray,
                    radius: float = 0.15, restitution: float = 1.0):
        pos = np.ascontiguousarray(pos, np.float32); vel = np.ascontiguousarray(vel, np.float32)
        count, n = pos.shape[0], pos.shape[1]
        st = prepare(statics)
        _check(self.L.pg_worlds_set_spheres(self.h, first, count, _ptr(st), len(st), n, _ptr(pos), _ptr(vel),
                                            np.float32(radius), np.float32(restitution)), "pg_worlds_set_spheres")
        self._sphere_n = n

    def _state_args(self, first, count, pos, vel, at_rest, what, writable):
        if first < 0 or count <= 0 or first + count > self.n_worlds:
            raise PgError(f"{what}: world range [{first}, {first + count}) outside 0..{self.n_worlds}")
        if getattr(self, "_sphere_n", None) is not None:      # sphere mode: every world holds the same count
            total = count * self._sphere_n
        else:
            total = sum(self.body_count(k, CLASS_DYNAMIC) for k in range(first, first + count))
        for a, per, dt_, name, opt in ((pos, 3, np.float32, "pos", False), (vel, 3, np.float32, "vel", False),
                                       (at_rest, 1, np.uint8, "at_rest", True)):
            if a is None and opt:
                continue
            if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dt_) or not a.flags["C_CONTIGUOUS"] or a.size != total * per \
                    or (writable and not a.flags["WRITEABLE"]):
                raise PgError(f"{what} {name}: need a C-contiguous {'writable ' if writable else ''}{np.dtype(dt_)} array of "
                              f"{total} x {per} items, got {type(a).__name__}"
                              + (f" dtype {a.dtype} shape {a.shape} contiguous {a.flags['C_CONTIGUOUS']}" if isinstance(a, np.ndarray) else ""))

    def upload_state(self, first: int, count: int, pos: np.ndarray, vel: np.ndarray, at_rest=None):
        self._state_args(first, count, pos, vel, at_rest, "upload_state", False)
        _check(self.L.pg_worlds_upload_state(self.h, first, count, _ptr(pos), _ptr(vel), _ptr(at_rest)),
               "pg_worlds_upload_state")

    def download_state(self, first: int, count: int, pos: np.ndarray, vel: np.ndarray, at_rest=None):
        self._state_args(first, count, pos, vel, at_rest, "download_state", True)
        _check(self.L.pg_worlds_download_state(self.h, first, count, _ptr(pos), _ptr(vel), _ptr(at_rest)),
               "pg_worlds_download_state")

    # -- stepping
    def step(self, dt: float, n_steps: int = 1, refresh_nodes: bool = True):
        _check(self.L.pg_worlds_step(self.h, np.float32(dt), n_steps, int(refresh_nodes)), "pg_worlds_step")

    def sync(self):
        _check(self.L.pg_worlds_sync(self.h), "pg_worlds_sync")

    def step_host(self, dt: float, n_steps: int, pos: np.ndarray, vel: np.ndarray, at_rest=None):
        """Host arrays in, n_steps steps, host arrays out (in place); chunk-pipelined inside the library."""
        self._state_args(0, self.n_worlds, pos, vel, at_rest, "step_host", True)
        _check(self.L.pg_worlds_step_host(self.h, np.float32(dt), n_steps, _ptr(pos), _ptr(vel), _ptr(at_rest)),
               "pg_worlds_step_host")

    def packed_dtype(self) -> np.dtype:
        """Record of one world in the packed host layout of step_host_packed (include/physics_gpu.h)."""
        n = self._sphere_n
        size = int(self.L.pg_worlds_packed_world_bytes(self.h))
        return np.dtype({"names": ["pos", "vel", "at_rest"], "formats": [("<f4", (n, 3)), ("<f4", (n, 3)), ("u1", (n,))],
                         "offsets": [0, 12 * n, 24 * n], "itemsize": size})

    def step_host_packed(self, dt: float, n_steps: int, state: np.ndarray):
        """`state`: C-contiguous array of n_worlds packed_dtype() records (ideally over pinned memory), stepped in place."""
        dtp = self.packed_dtype()
        if not isinstance(state, np.ndarray) or state.dtype != dtp or state.size != self.n_worlds or not state.flags["C_CONTIGUOUS"] \
                or not state.flags["WRITEABLE"]:
            raise PgError(f"step_host_packed: need a writable C-contiguous array of {self.n_worlds} packed_dtype() records")
        _check(self.L.pg_worlds_step_host_packed(self.h, np.float32(dt), n_steps, _ptr(state)), "pg_worlds_step_host_packed")

    def events(self) -> np.ndarray:
        over = C.c_int(0)
        n = _check(self.L.pg_worlds_events(self.h, None, 0, C.byref(over)), "pg_worlds_events")
        out = np.zeros(max(int(n), 1), dtype=PG_EVENT_DTYPE)
        n = _check(self.L.pg_worlds_events(self.h, _ptr(out), len(out), C.byref(over)), "pg_worlds_events")
        self.events_overflowed = bool(over.value)
        return out[:int(n)]

    def stats(self, dev_out_ptr: int | None = None, want_host: bool = True) -> dict | None:
        h = np.zeros(8, np.float64) if want_host else None
        _check(self.L.pg_worlds_stats(self.h, _ptr(h), C.c_void_p(dev_out_ptr) if dev_out_ptr else None),
               "pg_worlds_stats")
        return dict(zip(PG_STATS_FIELDS, h.tolist())) if want_host else None

    def stats_allreduce(self, comm: "Comm | None" = None, dev_out_ptr: int | None = None, want_host: bool = False) -> dict | None:
        """Statistics summed over every rank of `comm`, enqueued on the handle's stream (a collective)."""
        h = np.zeros(8, np.float64) if want_host else None
        _check(self.L.pg_worlds_stats_allreduce(self.h, comm.h if comm else None, _ptr(h),
                                                C.c_void_p(dev_out_ptr) if dev_out_ptr else None), "pg_worlds_stats_allreduce")
        return dict(zip(PG_STATS_FIELDS, h.tolist())) if want_host else None

    def debug_geometry(self, world: int = 0, device_ptr: int | None = None, max_records: int | None = None):
        """One PgDebugDraw record per body of `world` (players, statics, dynamics): what the reference's
        physics_debug_render reads from the bodies each frame (debug_renderer.c:11-213), built on the device.
        With `device_ptr` the kernel writes into that device buffer (a mapped GL buffer) and only the count
        comes back; otherwise a DEBUG_DTYPE array."""
        from .scenes import DEBUG_DTYPE
        if device_ptr is not None:
            if max_records is None or int(max_records) < 0:
                raise PgError("debug_geometry(device_ptr=...): max_records (the room of the device buffer, in records) is required")
            return _check(self.L.pg_worlds_debug_geometry(self.h, world, C.c_void_p(device_ptr), 1, int(max_records)),
                          "pg_worlds_debug_geometry")
        total = sum(self.body_count(world, c) for c in range(3))
        out = np.zeros(max(total, 1), dtype=DEBUG_DTYPE)
        n = _check(self.L.pg_worlds_debug_geometry(self.h, world, _ptr(out), 0, len(out)), "pg_worlds_debug_geometry")
        return out[:int(n)]

    @property
    def launches(self) -> int:
        return int(self.L.pg_worlds_launches(self.h))

    def last_step_ms(self) -> float:
        return float(self.L.pg_worlds_last_step_ms(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_worlds_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# -- per-pair entry points --------------------------------------------------------------------

def _pair_arrays(a, b=None):
    a = prepare(np.atleast_1d(a).copy())
    b = None if b is None else prepare(np.atleast_1d(b).copy())
    return a, b


def pair_max_move(a, t0, t1):
    a, _ = _pair_arrays(a)
    out = np.zeros(len(a), np.float32)
    _check(lib().pg_pair_max_move(_ptr(a), len(a), np.float32(t0), np.float32(t1), _ptr(out)), "pg_pair_max_move")
    return out


def pair_min_dist(a, b, t):
    a, b = _pair_arrays(a, b)
    t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float32), (len(a),)))
    out = np.zeros(len(a), np.float32)
    _check(lib().pg_pair_min_dist(_ptr(a), _ptr(b), len(a), _ptr(t), _ptr(out)), "pg_pair_min_dist")
    return out


def pair_interval_collision(a, b, t0, t1):
    a, b = _pair_arrays(a, b)
    hit = np.zeros(len(a), np.int32); ht = np.zeros(len(a), np.float32)
    _check(lib().pg_pair_interval_collision(_ptr(a), _ptr(b), len(a), np.float32(t0), np.float32(t1), _ptr(hit), _ptr(ht)),
           "pg_pair_interval_collision")
    return hit, ht


def pair_narrow_phase(a, b, dt):
    a, b = _pair_arrays(a, b)
    col = np.zeros(len(a), np.int32); ht = np.zeros(len(a), np.float32)
    _check(lib().pg_pair_narrow_phase(_ptr(a), _ptr(b), len(a), np.float32(dt), _ptr(col), _ptr(ht)), "pg_pair_narrow_phase")
    return col, ht


def pair_resolve(a, b, hit_time, dt):
    a, b = _pair_arrays(a, b)
    ht = np.ascontiguousarray(np.broadcast_to(np.asarray(hit_time, np.float32), (len(a),)))
    _check(lib().pg_pair_resolve(_ptr(a), _ptr(b), len(a), _ptr(ht), np.float32(dt)), "pg_pair_resolve")
    return a, b


def pair_visit(a, b, dt):
    a, b = _pair_arrays(a, b)
    ev = np.zeros(len(a), np.int32)
    _check(lib().pg_pair_visit(_ptr(a), _ptr(b), len(a), np.float32(dt), _ptr(ev)), "pg_pair_visit")
    return ev, a, b


class GpuLargeWorld:
    """One large world of dynamic spheres + static planes (PgLarge handle, uniform-grid broadphase)."""

    def __init__(self, max_spheres: int, max_static: int = 6, device: int = 0):
        self.L = lib()
        self.h = self.L.pg_large_create(max_spheres, max_static, device)
        if not self.h:
            raise PgError("pg_large_create failed: " + self.L.pg_last_error().decode())
        self.n = 0

    def set(self, statics: np.ndarray, pos: np.ndarray, vel: np.ndarray, radius=0.15, restitution: float = 1.0):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3); vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
        st = prepare(statics)
        rad_arr = None
        r_uniform = 0.0
        if np.ndim(radius) == 0:
            r_uniform = float(radius)
        else:
            rad_arr = np.ascontiguousarray(radius, np.float32)
        self.n = len(pos)
        _check(self.L.pg_large_set(self.h, _ptr(st), len(st), self.n, _ptr(pos), _ptr(vel), _ptr(rad_arr),
                                   np.float32(r_uniform), np.float32(restitution)), "pg_large_set")

    def set_players(self, players: np.ndarray):
        p = prepare(players)
        _check(self.L.pg_large_set_players(self.h, _ptr(p), len(p)), "pg_large_set_players")
        self.n_players = len(p)

    def get_players(self) -> np.ndarray:
        out = np.zeros(8, dtype=BODY_DTYPE)
        n = _check(self.L.pg_large_get_players(self.h, _ptr(out), len(out)), "pg_large_get_players")
        return out[:n]

    def upload_state(self, pos, vel, at_rest=None):
        _state_array(pos, (self.n, 3), np.float32, "upload_state pos")
        _state_array(vel, (self.n, 3), np.float32, "upload_state vel")
        _state_array(at_rest, (self.n,), np.uint8, "upload_state at_rest", optional=True)
        _check(self.L.pg_large_upload_state(self.h, _ptr(pos), _ptr(vel), _ptr(at_rest)), "pg_large_upload_state")

    def download_state(self, pos=None, vel=None, at_rest=None):
        pos = np.zeros((self.n, 3), np.float32) if pos is None else pos
        vel = np.zeros((self.n, 3), np.float32) if vel is None else vel
        at_rest = np.zeros(self.n, np.uint8) if at_rest is None else at_rest
        _state_array(pos, (self.n, 3), np.float32, "download_state pos", writable=True)
        _state_array(vel, (self.n, 3), np.float32, "download_state vel", writable=True)
        _state_array(at_rest, (self.n,), np.uint8, "download_state at_rest", writable=True)
        _check(self.L.pg_large_download_state(self.h, _ptr(pos), _ptr(vel), _ptr(at_rest)), "pg_large_download_state")
        return pos, vel, at_rest

    def step(self, dt: float, n_steps: int = 1):
        _check(self.L.pg_large_step(self.h, np.float32(dt), n_steps), "pg_large_step")

    def pairs(self, dt: float) -> np.ndarray:
        n = _check(self.L.pg_large_pairs(self.h, np.float32(dt), None, 0), "pg_large_pairs")
        out = np.zeros((max(int(n), 1), 2), np.int32)
        n = _check(self.L.pg_large_pairs(self.h, np.float32(dt), _ptr(out), len(out)), "pg_large_pairs")
        return out[:int(n)]

    def events(self) -> np.ndarray:
        over = C.c_int(0)
        n = _check(self.L.pg_large_events(self.h, None, 0, C.byref(over)), "pg_large_events")
        out = np.zeros(max(int(n), 1), dtype=PG_EVENT_DTYPE)
        n = _check(self.L.pg_large_events(self.h, _ptr(out), len(out), C.byref(over)), "pg_large_events")
        self.events_overflowed = bool(over.value)
        return out[:int(n)]

    def stats(self, dev_out_ptr: int | None = None, want_host: bool = True):
        h = np.zeros(8, np.float64) if want_host else None
        _check(self.L.pg_large_stats(self.h, _ptr(h), C.c_void_p(dev_out_ptr) if dev_out_ptr else None), "pg_large_stats")
        return dict(zip(PG_STATS_FIELDS, h.tolist())) if want_host else None

    def exactness(self) -> dict:
        it, pr = C.c_int(0), C.c_int(0)
        ca, hi = C.c_longlong(0), C.c_longlong(0)
        _check(self.L.pg_large_exactness(self.h, C.byref(it), C.byref(pr), C.byref(ca), C.byref(hi)), "pg_large_exactness")
        return {"iterations": it.value, "proven_exact": bool(pr.value), "candidates": ca.value, "hits": hi.value}

    def set_timing(self, on: bool):
        """Per-kernel CUDA-event timing (kernel_ms) is off by default; it costs an event pair per launch."""
        _check(self.L.pg_large_set_timing(self.h, int(on)), "pg_large_set_timing")

    def kernel_ms(self) -> dict:
        return {self.L.pg_large_kernel_name(k).decode(): float(self.L.pg_large_kernel_ms(self.h, k))
                for k in range(self.L.pg_large_kernel_count())}

    def last_step_ms(self) -> float:
        return float(self.L.pg_large_last_step_ms(self.h))

    @property
    def launches(self) -> int:
        return int(self.L.pg_large_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_large_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
