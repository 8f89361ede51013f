"""Seeded random bodies / worlds shared by the CPU and GPU parity tests."""
from __future__ import annotations

import numpy as np

from crux_b200.scenes import (BODY_DTYPE, COLLIDER_AABB, COLLIDER_CAPSULE, COLLIDER_PLANE, COLLIDER_SPHERE,
                              ENTITY_ITEM, ENTITY_PLAYER, ENTITY_WORLD, new_bodies)

LIVE_PAIRS = [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 3), (2, 3)]   # table entries with a body
DT = np.float32(0.01666)


def euler_xyz(a):
    """cglm glm_euler_xyz upper 3x3, float32, column-major flat (same op order as the shim)."""
    a = np.asarray(a, np.float32)
    sx, cx = np.sin(a[0]), np.cos(a[0]); sy, cy = np.sin(a[1]), np.cos(a[1]); sz, cz = np.sin(a[2]), np.cos(a[2])
    f = np.float32
    czsx, cxcz, sysz = f(cz * sx), f(cx * cz), f(sy * sz)
    return np.array([cy * cz, f(czsx * sy) + f(cx * sz), f(-cxcz * sy) + f(sx * sz),
                     -cy * sz, cxcz - f(sx * sysz), czsx + f(cx * sysz),
                     sy, -cy * sx, cx * cy], dtype=np.float32)


def node_xform(node_transform, pos, rot_deg, scale, parent=None):
    return node_transform(np.asarray(pos, np.float32), np.asarray(rot_deg, np.float32),
                          np.asarray(scale, np.float32), parent)


def rand_body(rng, t, dynamic, node_transform, near=None, entity=ENTITY_WORLD):
    """One random record of collider type t.  node_transform(pos, rot, scale, parent) -> 16 floats."""
    b = new_bodies(1)
    b["collider_type"] = t
    b["dynamic"] = int(dynamic)
    b["entity_type"] = entity
    b["restitution"] = rng.choice([0.0, 0.5, 0.8, 1.0])
    ctr = (rng.uniform(-2, 2, 3) if near is None else near + rng.normal(0, 0.6, 3)).astype(np.float32)
    b["position"] = ctr
    if dynamic:
        b["velocity"] = rng.uniform(-6, 6, 3).astype(np.float32) * rng.choice([0, 1, 1, 1])
    b["rotation"] = rng.choice([0, 0, 90, 37.5, -120], 3).astype(np.float32) if t in (0, 2) else 0
    b["scale"] = rng.choice([1.0, 0.5, 2.0]) if rng.random() < 0.7 else rng.uniform(0.5, 2, 3).astype(np.float32)
    if t == COLLIDER_AABB:
        b["shape"][0, :3] = rng.uniform(-0.2, 0.2, 3)
        b["shape"][0, 3:6] = rng.uniform(0.1, 1.0, 3)
    elif t == COLLIDER_SPHERE:
        b["shape"][0, :3] = rng.uniform(-0.1, 0.1, 3) * rng.choice([0, 1])
        b["shape"][0, 3] = rng.uniform(0.1, 0.6)
    elif t == COLLIDER_CAPSULE:
        b["shape"][0, :3] = [0, 0.25, 0]
        b["shape"][0, 3:6] = [0, 1.75, 0]
        b["shape"][0, 6] = 0.25
    else:
        n = rng.normal(0, 1, 3)
        n /= np.linalg.norm(n)
        if rng.random() < 0.5:
            n = np.eye(3)[rng.integers(3)] * rng.choice([-1, 1])
        b["shape"][0, :3] = n
        b["shape"][0, 3] = np.dot(n, ctr) + rng.normal(0, 0.4)
        b["position"] = 0
    b["has_node"] = 1
    b["has_parent"] = 1
    par = node_xform(node_transform, rng.uniform(-1, 1, 3) * rng.choice([0, 1]), rng.choice([0, 0, 90], 3), [1, 1, 1])
    b["parent_xform"] = par
    b["node_xform"] = node_xform(node_transform, b["position"][0], b["rotation"][0], b["scale"][0], par)
    return b


def rand_pairs(rng, n, node_transform):
    """n random (A, B) record pairs cycling through the 7 live collider pairs, A the lower type."""
    A = np.zeros(n, BODY_DTYPE)
    B = np.zeros(n, BODY_DTYPE)
    for k in range(n):
        ta, tb = LIVE_PAIRS[k % 7]
        a = rand_body(rng, ta, rng.random() < 0.8, node_transform)
        b = rand_body(rng, tb, (tb != 3) and rng.random() < 0.6, node_transform, near=a["position"][0])
        A[k], B[k] = a[0], b[0]
    return A, B


def mixed_world(rng, node_transform, n_static_boxes=4, n_dyn=24, with_player=True, L=6.0):
    """A small closed room: 6 planes + a few static AABBs; dynamic spheres and AABBs; one capsule player.

    Returns (statics, dynamics, players) records with scene nodes, like a loaded Crux scene.
    With a capsule player every dynamic body is an AABB: the reference's sphere-capsule and
    capsule-capsule table entries are empty functions (undefined behaviour upstream,
    distance.c:301-303, narrow_phase.c:522-524), so worlds that would visit such a pair are not
    valid parity inputs (scenes/bouncehouse.json is of this AABB-only kind).
    """
    from crux_b200.scenes import box_planes, plane_bodies
    ident = np.eye(4, dtype=np.float32).reshape(16)
    st = [plane_bodies(box_planes(L))]
    st[0]["has_node"] = 1
    st[0]["has_parent"] = 1
    for _ in range(n_static_boxes):
        b = new_bodies(1)
        b["collider_type"] = COLLIDER_AABB
        b["position"] = [rng.uniform(-L / 2 + 1, L / 2 - 1), rng.uniform(0.2, 1.0), rng.uniform(-L / 2 + 1, L / 2 - 1)]
        b["shape"][0, 3:6] = [rng.uniform(0.3, 0.8), rng.uniform(0.1, 0.4), rng.uniform(0.3, 0.8)]
        b["has_node"] = 1
        b["has_parent"] = 1
        b["node_xform"] = node_xform(node_transform, b["position"][0], [0, 0, 0], [1, 1, 1], ident)
        st.append(b)
    dy = []
    for k in range(n_dyn):
        b = new_bodies(1)
        b["dynamic"] = 1
        b["position"] = [rng.uniform(-L / 2 + 0.6, L / 2 - 0.6), rng.uniform(0.8, L - 0.8), rng.uniform(-L / 2 + 0.6, L / 2 - 0.6)]
        b["velocity"] = rng.uniform(-4, 4, 3)
        b["restitution"] = rng.choice([0.6, 0.8, 1.0])
        if k % 3 == 0 or with_player:
            b["collider_type"] = COLLIDER_AABB
            b["shape"][0, 3:6] = rng.uniform(0.1, 0.3, 3)
            if k % 6 == 0:
                b["entity_type"] = ENTITY_ITEM
        else:
            b["collider_type"] = COLLIDER_SPHERE
            b["shape"][0, 3] = rng.uniform(0.1, 0.3)
        b["has_node"] = 1
        b["has_parent"] = 1
        b["node_xform"] = node_xform(node_transform, b["position"][0], [0, 0, 0], [1, 1, 1], ident)
        dy.append(b)
    pl = []
    if with_player:
        b = new_bodies(1)
        b["collider_type"] = COLLIDER_CAPSULE
        b["entity_type"] = ENTITY_PLAYER
        b["shape"][0, :3] = [0, 0.25, 0]
        b["shape"][0, 3:6] = [0, 1.75, 0]
        b["shape"][0, 6] = 0.25
        b["position"] = [0.0, 0.5, 0.5]
        b["velocity"] = [1.0, 0.0, -0.5]
        b["has_node"] = 1
        b["has_parent"] = 1
        b["node_xform"] = node_xform(node_transform, b["position"][0], [0, 0, 0], [1, 1, 1], ident)
        pl.append(b)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, BODY_DTYPE)
    return cat(st), cat(dy), cat(pl)


def debug_gallery(rng, node_transform):
    """Bodies of every kind the reference's debug renderer draws (debug_renderer.c:11-213), never stepped:
    players  capsule with a scene node, capsule without one, AABB (node), sphere
    statics  planes, AABBs (node; rotated, scaled, under non-identity parents), spheres, a capsule
    dynamics AABBs (node), spheres with centre offsets and non-unit scales, a capsule (node)
    AABBs and dynamic capsules always carry a node: the reference dereferences it unconditionally."""
    def body(t, dynamic, node=True):
        b = rand_body(rng, t, dynamic, node_transform)
        if not node:
            b["has_node"] = 0; b["has_parent"] = 0
            b["node_xform"] = 0; b["parent_xform"] = 0
        return b
    pl = [body(COLLIDER_CAPSULE, False), body(COLLIDER_CAPSULE, False, node=False), body(COLLIDER_AABB, False),
          body(COLLIDER_SPHERE, False, node=False)]
    pl[1]["rotation"] = [12.5, -70.0, 33.0]; pl[1]["scale"] = [1.0, 1.5, 0.75]
    st = [body(COLLIDER_PLANE, False) for _ in range(2)] + [body(COLLIDER_AABB, False) for _ in range(6)] + \
         [body(COLLIDER_SPHERE, False, node=bool(k & 1)) for k in range(3)] + [body(COLLIDER_CAPSULE, False, node=False), body(COLLIDER_CAPSULE, False)]
    dy = [body(COLLIDER_AABB, True) for _ in range(5)] + [body(COLLIDER_SPHERE, True, node=bool(k & 1)) for k in range(6)] + \
         [body(COLLIDER_CAPSULE, True)]
    for group in (pl, st, dy):
        for b in group:
            b["entity_type"] = ENTITY_WORLD
    pl[0]["entity_type"] = ENTITY_PLAYER
    cat = lambda xs: np.concatenate(xs)
    return cat(st), cat(dy), cat(pl)


def debug_records_from_log(log, statics, dynamics, players):
    """Turn the reference's render log (ora.RefDebugLib.debug_render) into PgDebugDraw records, one per body in
    draw order (players, statics, dynamics), plus {record index: model} for capsules drawn without a scene node
    (those models are built on the host, pg_debug_capsule_model)."""
    from crux_b200.scenes import (DEBUG_DTYPE, DEBUG_AABB_LINES, DEBUG_SPHERE_MODEL, DEBUG_CAPSULE_MODEL,
                                  DEBUG_CAPSULE_HOST, CLASS_PLAYER, CLASS_STATIC, CLASS_DYNAMIC)
    MODEL, SUB, DRAW = 4, 3, 5
    order = [(CLASS_PLAYER, players), (CLASS_STATIC, statics), (CLASS_DYNAMIC, dynamics)]
    n = sum(len(a) for _, a in order)
    out = np.zeros(n, DEBUG_DTYPE)
    host_models = {}
    it = iter(log)
    k = 0
    for cls, arr in order:
        for i, b in enumerate(arr):
            out[k]["cls"] = cls; out[k]["index"] = i
            t = int(b["collider_type"])
            if t == COLLIDER_AABB:
                w, c, tag, m = next(it); assert w == MODEL and tag == 1 and np.array_equal(m, np.eye(4, dtype=np.float32).reshape(16))
                w, c, tag, v = next(it); assert w == SUB and c == 24
                w, c, tag, _ = next(it); assert w == DRAW and c == 24
                out[k]["kind"] = DEBUG_AABB_LINES; out[k]["n_indices"] = 24; out[k]["data"] = v
            elif t == COLLIDER_SPHERE:
                w, c, tag, m = next(it); assert w == MODEL and tag == 2
                w, c, tag, _ = next(it); assert w == DRAW and c == 2280
                out[k]["kind"] = DEBUG_SPHERE_MODEL; out[k]["n_indices"] = 2280; out[k]["data"][:16] = m
            elif t == COLLIDER_CAPSULE:
                w, c, tag, m = next(it); assert w == MODEL and tag == 2
                w, c, tag, _ = next(it); assert w == DRAW and c == 4920
                out[k]["n_indices"] = 4920
                through_node = bool(b["has_node"]) and cls != CLASS_STATIC
                if through_node:
                    out[k]["kind"] = DEBUG_CAPSULE_MODEL; out[k]["data"][:16] = m
                else:
                    out[k]["kind"] = DEBUG_CAPSULE_HOST; host_models[k] = m.copy()
            k += 1
    assert next(it, None) is None, "log entries left over"
    return out, host_models


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same_bits(x, y):
    """Bit equality of float arrays, treating any-NaN == any-NaN."""
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    return bool(np.all((x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))))


def state_equal(a, b):
    return (same_bits(a["position"], b["position"]) and same_bits(a["velocity"], b["velocity"])
            and np.array_equal(a["at_rest"] != 0, b["at_rest"] != 0))
