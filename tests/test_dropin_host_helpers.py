"""The host-side public symbols of the drop-in library (libcrux_physics.so: crux_colliders.c) against the
UNMODIFIED reference's own functions (oracle/_ref/libcrux_ref.so is aabb.c, sphere.c, physics/utils.c,
narrow_phase.c ... compiled where they lie): AABB_merge, AABB_update, AABB_update_by_vertex,
AABB_intersect_AABB, AABB_intersect_plane (aabb.c:8-121), sphere_intersect_sphere (sphere.c:5-12),
distance_point_to_segment (physics/utils.c:33-60), ray_intersect_AABB (narrow_phase.c:702-739),
get_collision_behavior (collider.c:7-17) -- SURVEY.md rows a33-a36.  Same structs, same calls, random and edge
inputs, results compared bit for bit.  Also runs the reference's own Unity test (test/physics/test_aabb.c)
linked against the drop-in library.  No GPU needed: these are host functions upstream and here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "crux_b200", "lib", "libcrux_physics.so")
STUBS_SO = os.path.join(ROOT, "oracle", "_build", "libdropin_stubs.so")
UNITY_DROPIN = os.path.join(ROOT, "oracle", "_ref", "test_aabb_dropin")


class AABB(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("extents", C.c_float * 3), ("initialized", C.c_bool)]


class Plane(C.Structure):
    _fields_ = [("normal", C.c_float * 3), ("distance", C.c_float)]


class Sphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float)]


V3 = C.c_float * 3
M3 = C.c_float * 9


def _bind(dll):
    dll.AABB_merge.argtypes = [C.POINTER(AABB), C.POINTER(AABB)]
    dll.AABB_update.argtypes = [C.POINTER(AABB), M3, V3, V3, C.POINTER(AABB)]
    dll.AABB_update_by_vertex.argtypes = [C.POINTER(AABB), V3]
    dll.AABB_intersect_AABB.argtypes = [C.POINTER(AABB), C.POINTER(AABB)]; dll.AABB_intersect_AABB.restype = C.c_bool
    dll.AABB_intersect_plane.argtypes = [C.POINTER(AABB), C.POINTER(Plane)]; dll.AABB_intersect_plane.restype = C.c_bool
    dll.sphere_intersect_sphere.argtypes = [C.POINTER(Sphere), C.POINTER(Sphere)]; dll.sphere_intersect_sphere.restype = C.c_bool
    dll.distance_point_to_segment.argtypes = [V3, V3, V3]; dll.distance_point_to_segment.restype = C.c_float
    dll.ray_intersect_AABB.argtypes = [V3, V3, C.POINTER(AABB), C.POINTER(C.c_float), C.c_float]; dll.ray_intersect_AABB.restype = C.c_bool
    dll.get_collision_behavior.argtypes = [C.c_int, C.c_int]; dll.get_collision_behavior.restype = C.c_int
    return dll


@pytest.fixture(scope="module")
def libs(ref_lib):
    assert C.sizeof(AABB) == 28
    if not os.path.exists(HOST_SO):
        pytest.fail("libcrux_physics.so not built (python -m crux_b200.build)")
    if not os.path.exists(STUBS_SO):
        ora.build_oracles()
    C.CDLL(STUBS_SO, mode=C.RTLD_GLOBAL)
    return _bind(C.CDLL(HOST_SO)), _bind(C.CDLL(ora.REF_SO))


def _box(rng, init=True):
    b = AABB()
    b.center[:] = rng.uniform(-3, 3, 3).astype(np.float32).tolist()
    b.extents[:] = rng.uniform(0.05, 2, 3).astype(np.float32).tolist()
    b.initialized = init
    return b


def _bits(x):
    return np.asarray(list(x), np.float32).view(np.uint32).tolist()


def _same_box(a, b):
    return _bits(a.center) == _bits(b.center) and _bits(a.extents) == _bits(b.extents) and bool(a.initialized) == bool(b.initialized)


def _copy(b):
    c = type(b)()
    C.memmove(C.byref(c), C.byref(b), C.sizeof(b))
    return c


def test_aabb_helpers_equal_reference(libs):
    ours, ref = libs
    rng = np.random.default_rng(20261020)
    n_true = 0
    for k in range(2000):
        a, b = _box(rng, init=(k % 7 != 0)), _box(rng)
        if k % 5 == 0:                                  # touching faces: `>` makes touching an intersection (aabb.c:100-108)
            b.center[0] = np.float32(a.center[0]) + np.float32(a.extents[0]) + np.float32(b.extents[0])
            b.center[1] = a.center[1]; b.center[2] = a.center[2]
        r1, r2 = ours.AABB_intersect_AABB(C.byref(a), C.byref(b)), ref.AABB_intersect_AABB(C.byref(a), C.byref(b))
        assert r1 == r2, k
        n_true += int(r1)
        a1, a2 = _copy(a), _copy(a)
        ours.AABB_merge(C.byref(a1), C.byref(b)); ref.AABB_merge(C.byref(a2), C.byref(b))
        assert _same_box(a1, a2), ("merge", k)
        v = V3(*rng.uniform(-4, 4, 3).astype(np.float32).tolist())
        a1, a2 = _copy(a), _copy(a)
        ours.AABB_update_by_vertex(C.byref(a1), v); ref.AABB_update_by_vertex(C.byref(a2), v)
        assert _same_box(a1, a2), ("by_vertex", k)
        pl = Plane()
        nrm = rng.normal(0, 1, 3); nrm /= np.linalg.norm(nrm)
        if k % 3 == 0:
            nrm = np.eye(3)[k % 3 - 0] * (1 if k % 2 else -1)
        pl.normal[:] = nrm.astype(np.float32).tolist()
        pl.distance = float(np.float32(np.dot(nrm, list(a.center)) + rng.normal(0, 1.0)))
        assert ours.AABB_intersect_plane(C.byref(a), C.byref(pl)) == ref.AABB_intersect_plane(C.byref(a), C.byref(pl)), ("plane", k)
        # AABB_update: rotation from a random orthonormal-ish matrix (also scaled / sheared ones: no assumption upstream)
        q, _ = np.linalg.qr(rng.normal(0, 1, (3, 3)))
        if k % 4 == 0:
            q = q * rng.uniform(0.5, 2.0)
        rot = M3(*q.astype(np.float32).reshape(9).tolist())
        tr = V3(*rng.uniform(-5, 5, 3).astype(np.float32).tolist())
        sc = V3(*rng.choice([1.0, 0.5, 2.0, -1.5], 3).astype(np.float32).tolist())
        d1, d2 = AABB(), AABB()
        ours.AABB_update(C.byref(a), rot, tr, sc, C.byref(d1)); ref.AABB_update(C.byref(a), rot, tr, sc, C.byref(d2))
        assert _bits(d1.center) == _bits(d2.center) and _bits(d1.extents) == _bits(d2.extents), ("update", k)
    assert 100 < n_true < 1900


def test_sphere_segment_ray_helpers_equal_reference(libs):
    ours, ref = libs
    rng = np.random.default_rng(7)
    hits = 0
    for k in range(3000):
        s1, s2 = Sphere(), Sphere()
        s1.center[:] = rng.uniform(-2, 2, 3).astype(np.float32).tolist(); s1.radius = float(np.float32(rng.uniform(0.1, 1.5)))
        s2.center[:] = rng.uniform(-2, 2, 3).astype(np.float32).tolist(); s2.radius = float(np.float32(rng.uniform(0.1, 1.5)))
        assert ours.sphere_intersect_sphere(C.byref(s1), C.byref(s2)) == ref.sphere_intersect_sphere(C.byref(s1), C.byref(s2)), k
        p, a, b = (V3(*rng.uniform(-3, 3, 3).astype(np.float32).tolist()) for _ in range(3))
        if k % 11 == 0:
            b = V3(*list(a))                                 # degenerate segment
        d1, d2 = ours.distance_point_to_segment(p, a, b), ref.distance_point_to_segment(p, a, b)
        assert np.float32(d1).view(np.uint32) == np.float32(d2).view(np.uint32) or (d1 != d1 and d2 != d2), ("segment", k, d1, d2)
        box = _box(rng)
        org = V3(*rng.uniform(-5, 5, 3).astype(np.float32).tolist())
        dirv = rng.uniform(-6, 6, 3).astype(np.float32)
        if k % 4 == 0:
            dirv[rng.integers(3)] = 0.0                      # axis-parallel ray: the fabsf(d) < EPSILON slab branch
        if k % 9 == 0:
            org = V3(*[float(np.float32(c + rng.uniform(-0.5, 0.5) * e)) for c, e in zip(box.center, box.extents)])   # starts inside
        dv = V3(*dirv.tolist())
        end = float(np.float32(rng.choice([0.01666, 0.5, 2.0])))
        t1, t2 = C.c_float(-7.0), C.c_float(-7.0)
        h1, h2 = ours.ray_intersect_AABB(org, dv, C.byref(box), C.byref(t1), end), ref.ray_intersect_AABB(org, dv, C.byref(box), C.byref(t2), end)
        assert h1 == h2, ("ray", k)
        assert np.float32(t1.value).view(np.uint32) == np.float32(t2.value).view(np.uint32), ("ray time", k, t1.value, t2.value)
        hits += int(h1)
    assert hits > 100
    for a in range(4):
        for b in range(4):
            assert ours.get_collision_behavior(a, b) == ref.get_collision_behavior(a, b)


def test_reference_unity_suite_against_the_dropin_library():
    """test/physics/test_aabb.c (the reference's whole physics test suite), compiled from where it lies and linked
    against libcrux_physics.so instead of the reference's aabb.c: 5 tests, 0 failures."""
    if not os.path.exists(UNITY_DROPIN):
        if os.path.isdir(ora.REFERENCE_ROOT):
            ora.build_oracles()
        if not os.path.exists(UNITY_DROPIN):
            pytest.skip("oracle/_ref/test_aabb_dropin not built and /root/reference absent")
    r = subprocess.run([UNITY_DROPIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "5 Tests 0 Failures 0 Ignored" in r.stdout
    # and it really is the drop-in's code that ran
    ldd = subprocess.run(["ldd", UNITY_DROPIN], capture_output=True, text=True).stdout
    assert "libcrux_physics.so" in ldd and "crux_b200/lib" in ldd
