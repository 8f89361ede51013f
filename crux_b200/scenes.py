"""Synthetic scenes and body records for the Crux physics step.

The synthetic inputs are the ones SURVEY.md section 8(d) fixes so that builder, tests and judge agree:

* Box: six static planes (n.x = d), laid out like /root/reference/scenes/scene_sphere.json:19-104
  scaled to edge L: floor (0,1,0),0; ceiling (0,-1,0),-L; (+-1,0,0),-L/2; (0,0,+-1),-L/2.
* Spheres: local centre 0, radius 0.15, scale 1, restitution 1, dynamic, entity type WORLD.
* RNG: 32-bit LCG s = s*1664525 + 1013904223, u = (s >> 8) / 2**24, seed 12345 + world index,
  six draws per body in order x, y, z, vx, vy, vz with x,z = (u-0.5)(L-1), y = 0.5 + u(L-1),
  v = (u-0.5)*8, all in binary32.

Records: BODY_DTYPE is the flat, pointer-free body record of include/physics_gpu.h (`PgBody`),
field for field the reference's `struct PhysicsBody` (/root/reference/include/physics/world.h:8-26)
plus the scene-node matrices the physics path dereferences.
"""
from __future__ import annotations

import numpy as np

COLLIDER_AABB, COLLIDER_SPHERE, COLLIDER_CAPSULE, COLLIDER_PLANE = 0, 1, 2, 3
ENTITY_GROUPING, ENTITY_WORLD, ENTITY_ITEM, ENTITY_PLAYER = 0, 1, 2, 3
CLASS_STATIC, CLASS_DYNAMIC, CLASS_PLAYER = 0, 1, 2

MAX_DELTA_TIME = np.float32(0.01666)          # world.c:18 (double literal, assigned to a float)
DT_60HZ = np.float32(1.0 / 60.0)              # what the frame loop feeds; clamped to 0.01666f

_CORE_FIELDS = [
    ("position", "<f4", (3,)),
    ("rotation", "<f4", (3,)),
    ("scale", "<f4", (3,)),
    ("velocity", "<f4", (3,)),
    ("restitution", "<f4"),
    ("shape", "<f4", (7,)),
    ("collider_type", "<i4"),
    ("entity_type", "<i4"),
    ("dynamic", "<i4"),
    ("at_rest", "<i4"),
    ("has_node", "<i4"),
    ("has_parent", "<i4"),
    ("node_xform", "<f4", (16,)),
    ("parent_xform", "<f4", (16,)),
]
# PgBody (include/physics_gpu.h): the core record followed by two host-computed rotation matrices.
BODY_DTYPE = np.dtype(_CORE_FIELDS + [("euler_raw", "<f4", (9,)), ("euler_node", "<f4", (9,))])
CORE_DTYPE = np.dtype(_CORE_FIELDS)            # == oracle/oracle_abi.h OraBody (232 bytes)
EVENT_DTYPE = np.dtype([("step", "<i4"), ("event_type", "<i4"), ("a_class", "<i4"),
                        ("a_index", "<i4"), ("b_class", "<i4"), ("b_index", "<i4")])
assert CORE_DTYPE.itemsize == 232 and BODY_DTYPE.itemsize == 304
# PgDebugDraw (include/physics_gpu.h): one debug-geometry record per body (src/physics/debug_renderer.c)
DEBUG_DTYPE = np.dtype([("kind", "<i4"), ("cls", "<i4"), ("index", "<i4"), ("n_indices", "<i4"), ("data", "<f4", (24,))])
DEBUG_NONE, DEBUG_AABB_LINES, DEBUG_SPHERE_MODEL, DEBUG_CAPSULE_MODEL, DEBUG_CAPSULE_HOST = 0, 1, 2, 3, 4
assert DEBUG_DTYPE.itemsize == 112

_IDENTITY16 = np.eye(4, dtype=np.float32).reshape(16)

LCG_A = 1664525
LCG_C = 1013904223
_MASK = np.uint64(0xFFFFFFFF)


def _lcg_tables(n: int):
    """A_k, C_k (k = 1..n) with s_k = A_k * s_0 + C_k (mod 2**32), built by doubling."""
    A = np.empty(n + 1, dtype=np.uint64)
    C = np.empty(n + 1, dtype=np.uint64)
    A[0], C[0] = 1, 0
    if n >= 1:
        A[1], C[1] = LCG_A, LCG_C
    m = 1
    while m < n:
        hi = min(2 * m, n)
        k = np.arange(m + 1, hi + 1)
        A[k] = (A[k - m] * A[m]) & _MASK
        C[k] = ((A[k - m] * C[m]) & _MASK) + C[k - m] & _MASK
        m = hi
    return A[1:], C[1:]


def lcg_uniform(seeds: np.ndarray, n_draws: int) -> np.ndarray:
    """u[w, k] in [0,1): the k-th draw (k = 0..n_draws-1) of the LCG seeded with seeds[w]. float32."""
    seeds = np.asarray(seeds, dtype=np.uint64).reshape(-1, 1)
    A, C = _lcg_tables(n_draws)
    s = ((seeds * A[None, :]) & _MASK) + C[None, :] & _MASK
    return ((s >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def box_planes(L: float) -> np.ndarray:
    """(6, 4) float32: normal xyz, distance -- the closed box of SURVEY.md section 8(d)."""
    h = np.float32(L) / np.float32(2)
    return np.array([
        [0, 1, 0, 0],
        [0, -1, 0, -np.float32(L)],
        [1, 0, 0, -h],
        [-1, 0, 0, -h],
        [0, 0, 1, -h],
        [0, 0, -1, -h],
    ], dtype=np.float32)


def sphere_worlds(n_worlds: int, n_spheres: int, L: float, first_world: int = 0,
                  radius: float = 0.15, speed: float = 8.0):
    """Positions and velocities of `n_worlds` independent worlds of `n_spheres` spheres each.

    Returns (pos, vel) float32 arrays of shape (n_worlds, n_spheres, 3).  World w uses seed
    12345 + first_world + w, so a rank that owns worlds [a, b) generates exactly its shard.
    """
    seeds = 12345 + first_world + np.arange(n_worlds, dtype=np.uint64)
    u = lcg_uniform(seeds, 6 * n_spheres).reshape(n_worlds, n_spheres, 6)
    Lm1 = np.float32(L - 1.0)
    half = np.float32(0.5)
    pos = np.empty((n_worlds, n_spheres, 3), dtype=np.float32)
    pos[..., 0] = (u[..., 0] - half) * Lm1
    pos[..., 1] = half + u[..., 1] * Lm1
    pos[..., 2] = (u[..., 2] - half) * Lm1
    vel = ((u[..., 3:6] - half) * np.float32(speed)).astype(np.float32)
    return pos, vel


def new_bodies(n: int) -> np.ndarray:
    b = np.zeros(n, dtype=BODY_DTYPE)
    b["scale"] = 1.0
    b["entity_type"] = ENTITY_WORLD
    b["node_xform"] = _IDENTITY16
    b["parent_xform"] = _IDENTITY16
    b["euler_raw"] = np.eye(3, dtype=np.float32).reshape(9)
    b["euler_node"] = np.eye(3, dtype=np.float32).reshape(9)
    return b


def plane_bodies(planes: np.ndarray) -> np.ndarray:
    b = new_bodies(len(planes))
    b["collider_type"] = COLLIDER_PLANE
    b["shape"][:, 0:4] = planes
    return b


def sphere_bodies(pos: np.ndarray, vel: np.ndarray, radius: float = 0.15,
                  restitution: float = 1.0) -> np.ndarray:
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    vel = np.asarray(vel, dtype=np.float32).reshape(-1, 3)
    b = new_bodies(len(pos))
    b["collider_type"] = COLLIDER_SPHERE
    b["dynamic"] = 1
    b["position"] = pos
    b["velocity"] = vel
    b["restitution"] = restitution
    b["shape"][:, 3] = radius
    return b


def bouncehouse_world(world_index: int = 0, n_spheres: int = 256, L: float = 10.0):
    """(static_bodies, dynamic_bodies) records of one synthetic 'bouncehouse' world (config C3)."""
    pos, vel = sphere_worlds(1, n_spheres, L, first_world=world_index)
    return plane_bodies(box_planes(L)), sphere_bodies(pos[0], vel[0])


# config table: SURVEY.md section 8(d) "Sizes"
CONFIGS = {
    "c3": dict(n_worlds=4096, n_spheres=256, L=10.0),
    "c2": dict(n_worlds=1, n_spheres=65536, L=64.0),
    "c4": dict(n_worlds=1, n_spheres=4194304, L=256.0),
}


def translate16(pos) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = np.asarray(pos, np.float32)        # cglm column-major: translation in column 3
    return m.reshape(16)


def level_slabs(L: float, spacing: float = 8.0, extents=(4.0, 0.25, 4.0)) -> np.ndarray:
    """Static level geometry of config C5: axis-aligned AABB slabs on an
    `spacing`-unit lattice inside the L-box, checkerboard so that spheres can pass between floors;
    identity-rotation scene nodes under an identity parent."""
    n = int(L // spacing)
    out = []
    for j in range(1, n):
        for i in range(n):
            for k in range(n):
                if (i + j + k) % 2:
                    continue
                out.append((-L / 2 + spacing * (i + 0.5), spacing * j, -L / 2 + spacing * (k + 0.5)))
    b = new_bodies(len(out))
    b["collider_type"] = COLLIDER_AABB
    b["position"] = np.asarray(out, np.float32).reshape(-1, 3)
    b["shape"][:, 3:6] = np.asarray(extents, np.float32)
    b["has_node"] = 1
    b["has_parent"] = 1
    for r in b:
        r["node_xform"] = translate16(r["position"])
    return b


def player_capsule(pos=(0.0, 0.0, 2.0), rot=(0.0, 180.0, 0.0), vel=(0.0, 0.0, 0.0)) -> np.ndarray:
    """The player body scene_load creates (scene.c:303-332): capsule A=(0,.25,0) B=(0,1.75,0) r=.25."""
    b = new_bodies(1)
    b["collider_type"] = COLLIDER_CAPSULE
    b["entity_type"] = ENTITY_PLAYER
    b["shape"][0, :3] = [0.0, 0.25, 0.0]
    b["shape"][0, 3:6] = [0.0, 1.75, 0.0]
    b["shape"][0, 6] = 0.25
    b["position"] = np.asarray(pos, np.float32)
    b["rotation"] = np.asarray(rot, np.float32)
    b["velocity"] = np.asarray(vel, np.float32)
    b["has_node"] = 1
    b["has_parent"] = 1
    return b


def level_world(n_spheres: int, L: float, seed_world: int = 0, with_player: bool = True):
    """Config C5: (statics, sphere pos, sphere vel, players).  Statics = the six box planes followed
    by the slab level; spheres as in C2."""
    planes = plane_bodies(box_planes(L))
    planes["has_node"] = 1
    planes["has_parent"] = 1
    statics = np.concatenate([planes, level_slabs(L)])
    pos, vel = sphere_worlds(1, n_spheres, L, first_world=seed_world)
    players = player_capsule() if with_player else np.zeros(0, dtype=BODY_DTYPE)
    return statics, pos[0], vel[0], players


def sphere_bodies_with_nodes(pos, vel, radius: float = 0.15, restitution: float = 1.0) -> np.ndarray:
    """Dynamic spheres carrying a scene node (translation = position), as a loaded scene has them."""
    b = sphere_bodies(pos, vel, radius, restitution)
    b["has_node"] = 1
    m = np.tile(np.eye(4, dtype=np.float32).reshape(16), (len(b), 1))
    m[:, 12:15] = b["position"]
    b["node_xform"] = m
    return b


CONFIGS["c5"] = dict(n_worlds=1, n_spheres=1048576, L=160.0)
