"""The plain-C restatement (oracle/crux_oracle.c) against committed outputs of the UNMODIFIED
reference (tests/golden/, written by tests/make_golden.py) and against the reference's own Unity
vectors.  Runs on any box with gcc; no GPU, no /root/reference."""
import ctypes as C
import os

import numpy as np
import pytest

import gen
import ora
from crux_b200 import scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_pair_known_answers(oracle_lib):
    z = np.load(os.path.join(G, "pairs_kat.npz"))
    A, B, t = z["A"], z["B"], z["t"]
    for k in range(len(A)):
        a, b = A[k:k + 1], B[k:k + 1]
        assert gen.same_bits(oracle_lib.k_min_dist(a, b, t[k]), z["dist"][k]), k
        assert gen.same_bits(oracle_lib.k_max_move(a, 0.0, gen.DT), z["maxmove"][k]), k
        assert oracle_lib.k_interval(a, b, 0.0, gen.DT)[0] == z["interval"][k], k
        c, ht = oracle_lib.k_narrow(a, b, gen.DT)
        assert c == z["colliding"][k] and gen.same_bits(ht, z["hit_time"][k]), k
        ev, aa, bb = oracle_lib.k_visit(a, b, gen.DT)
        assert ev == z["visit"][k], k
        assert gen.state_equal(aa, z["A_after"][k:k + 1]) and gen.state_equal(bb, z["B_after"][k:k + 1]), k
        _, aa, bb = oracle_lib.k_resolve(a, b, z["resolve_time"][k], gen.DT)
        assert gen.state_equal(aa, z["A_resolved"][k:k + 1]) and gen.state_equal(bb, z["B_resolved"][k:k + 1]), k


def test_sphere_world_trajectory(oracle_lib):
    z = np.load(os.path.join(G, "spheres_256.npz"))
    st, dy = scenes.bouncehouse_world(0)
    w = ora.World(oracle_lib, st, dy)
    done = 0
    for target in (1, 10, 100, 1000):
        w.step(scenes.DT_60HZ, target - done)
        done = target
        assert gen.state_equal(w.read(1), z[f"state_{target}"]), target
        if target == 100:
            assert np.array_equal(w.events(clear=False), z["events_first100"])
    assert len(w.events()) == int(z["n_events_1000"]) == 13725      # the survey's probe count


def test_mixed_world_with_scene_nodes(oracle_lib):
    z = np.load(os.path.join(G, "mixed_world.npz"))
    w = ora.World(oracle_lib, z["statics"], z["dynamics"], z["players"])
    done = 0
    for target in (25, 50, 150):
        w.step(scenes.DT_60HZ, target - done, 1)
        done = target
        for cls, name in ((0, "S"), (1, "D"), (2, "P")):
            got, want = w.read(cls), z[f"{name}_{target}"]
            assert gen.state_equal(got, want), (target, name)
            assert gen.same_bits(got["node_xform"], want["node_xform"]), (target, name)
    assert np.array_equal(w.events(), z["events"])


def test_reference_unity_vectors(oracle_lib):
    """test/physics/test_aabb.c:24-141 restated on the port's AABB helpers."""
    z = np.load(os.path.join(G, "aabb_unity.npz"))
    d = oracle_lib.dll
    d.ora_aabb_intersect_aabb.restype = C.c_int
    d.ora_aabb_intersect_plane.restype = C.c_int
    assert d.ora_aabb_intersect_aabb(_p(z["aabb_true_a"]), _p(z["aabb_true_b"])) == 1
    assert d.ora_aabb_intersect_aabb(_p(z["aabb_false_a"]), _p(z["aabb_false_b"])) == 0
    assert d.ora_aabb_intersect_plane(_p(z["plane_box"]), _p(z["plane_true_1"])) == 1
    assert d.ora_aabb_intersect_plane(_p(z["plane_box"]), _p(z["plane_true_2"])) == 1
    assert d.ora_aabb_intersect_plane(_p(z["plane_box"]), _p(z["plane_false_1"])) == 0
    assert d.ora_aabb_intersect_plane(_p(z["plane_box"]), _p(z["plane_false_2"])) == 0
    # AABB_update under glm_rotate_y(identity, rad(90)): rotation picked from the 4x4
    a = np.float32(np.deg2rad(np.float32(90.0)))
    c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
    rot = np.array([c, 0, -s, 0, 1, 0, s, 0, c], np.float32)          # column-major 3x3 of glm_rotate_y
    out = np.zeros(6, np.float32)
    d.ora_aabb_update(_p(z["update_src"]), _p(rot), _p(z["update_translation"]), _p(z["update_scale"]), _p(out))
    assert np.allclose(out, z["update_expected"], atol=float(z["update_tol"]))


def test_helper_functions_outside_the_step(oracle_lib):
    """AABB_merge / AABB_update_by_vertex / sphere_intersect_sphere / distance_point_to_segment /
    behaviour and event tables."""
    d = oracle_lib.dll
    a = np.array([0, 0, 0, 1, 1, 1], np.float32); b = np.array([2, 0, 0, 1, 2, 1], np.float32)
    ai = C.c_int(1)
    d.ora_aabb_merge(_p(a), C.byref(ai), _p(b), 1)
    assert np.allclose(a, [1, 0, 0, 2, 2, 1])
    e = np.zeros(6, np.float32); ei = C.c_int(0)
    d.ora_aabb_merge(_p(e), C.byref(ei), _p(b), 1)
    assert np.array_equal(e, b) and ei.value == 1
    box = np.array([0, 0, 0, 1, 1, 1], np.float32)
    d.ora_aabb_update_by_vertex(_p(box), _p(np.array([3, 0, 0], np.float32)))
    assert np.allclose(box, [1, 0, 0, 2, 1, 1])
    d.ora_sphere_intersect_sphere.restype = C.c_int
    assert d.ora_sphere_intersect_sphere(_p(np.array([0, 0, 0, 1], np.float32)), _p(np.array([2, 0, 0, 1], np.float32))) == 1
    assert d.ora_sphere_intersect_sphere(_p(np.array([0, 0, 0, 1], np.float32)), _p(np.array([2.1, 0, 0, 1], np.float32))) == 0
    d.ora_distance_point_to_segment.restype = C.c_float
    f = lambda p, s0, s1: d.ora_distance_point_to_segment(_p(np.array(p, np.float32)), _p(np.array(s0, np.float32)), _p(np.array(s1, np.float32)))
    assert abs(f([0, 1, 0], [-1, 0, 0], [1, 0, 0]) - 1.0) < 1e-6
    assert abs(f([3, 0, 0], [-1, 0, 0], [1, 0, 0]) - 2.0) < 1e-6
    d.ora_collision_behavior.restype = C.c_int
    d.ora_event_type.restype = C.c_int
    W, I, P, Gp = scenes.ENTITY_WORLD, scenes.ENTITY_ITEM, scenes.ENTITY_PLAYER, scenes.ENTITY_GROUPING
    assert d.ora_collision_behavior(W, W) == 0 and d.ora_collision_behavior(I, P) == 1 and d.ora_collision_behavior(Gp, W) == 1
    assert d.ora_collision_behavior(P, W) == 0 and d.ora_collision_behavior(P, P) == 0
    assert d.ora_event_type(I, P) == 1 and d.ora_event_type(P, I) == 1 and d.ora_event_type(W, P) == 0


def test_remove_body_swaps_last(oracle_lib):
    """physics_remove_body (world.c:147-172): swap-with-last changes the solve order."""
    st, dy = scenes.bouncehouse_world(1, n_spheres=8)
    w = ora.World(oracle_lib, st, dy)
    before = w.read(1)
    w.remove(1, 2)
    after = w.read(1)
    assert len(after) == 7 and gen.state_equal(after[2:3], before[7:8]) and gen.state_equal(after[:2], before[:2])
