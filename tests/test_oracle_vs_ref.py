"""The plain-C restatement against the UNMODIFIED reference compiled in this container
(oracle/_ref/libcrux_ref.so).  Skipped on boxes with neither /root/reference nor a prebuilt _ref;
tests/test_oracle_golden.py covers those through committed fixtures."""
import subprocess

import numpy as np
import pytest

import gen
import ora
from crux_b200 import scenes


def _nt(orc):
    from make_golden import oracle_node_transform
    return oracle_node_transform(orc)


def test_reference_own_unity_test_passes(ref_lib):
    import os
    exe = os.path.join(ora.ORACLE_DIR, "_ref", "test_aabb")
    if not os.path.exists(exe):
        pytest.skip("test_aabb not built")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "5 Tests 0 Failures" in out.stdout


def test_struct_sizes_match_survey(ref_lib):
    import ctypes as C
    sizes = (C.c_int * 7)()
    ref_lib.dll.ref_sizes(sizes)
    body, collider, result, entity, _node, event, aabb = list(sizes)
    assert (body, collider, result, entity, event, aabb) == (120, 32, 24, 104, 64, 28)


def test_random_pairs_all_tables(ref_lib, oracle_lib):
    rng = np.random.default_rng(99)
    A, B = gen.rand_pairs(rng, 1400, _nt(oracle_lib))
    fired = 0
    for k in range(len(A)):
        a, b = A[k:k + 1], B[k:k + 1]
        t = np.float32(rng.uniform(0, gen.DT))
        assert gen.same_bits(ref_lib.k_min_dist(a, b, t), oracle_lib.k_min_dist(a, b, t)), k
        assert ref_lib.k_interval(a, b, 0, gen.DT)[0] == oracle_lib.k_interval(a, b, 0, gen.DT)[0], k
        (c1, h1), (c2, h2) = ref_lib.k_narrow(a, b, gen.DT), oracle_lib.k_narrow(a, b, gen.DT)
        assert c1 == c2 and gen.same_bits(h1, h2), k
        (e1, a1, b1), (e2, a2, b2) = ref_lib.k_visit(a, b, gen.DT), oracle_lib.k_visit(a, b, gen.DT)
        assert e1 == e2 and gen.state_equal(a1, a2) and gen.state_equal(b1, b2), k
        fired += e1
    assert fired > 200


@pytest.mark.parametrize("world_index", [0, 5])
def test_sphere_world_300_steps(ref_lib, oracle_lib, world_index):
    st, dy = scenes.bouncehouse_world(world_index)
    r, o = ora.World(ref_lib, st, dy), ora.World(oracle_lib, st, dy)
    g = ora.World(oracle_lib, st, dy)
    r.step(scenes.DT_60HZ, 300); o.step(scenes.DT_60HZ, 300)
    oracle_lib.dll.ora_world_step_grid(g.h, scenes.DT_60HZ, 300, 1)
    assert gen.state_equal(r.read(1), o.read(1)) and gen.state_equal(r.read(1), g.read(1))
    er = r.events()
    assert np.array_equal(er, o.events()) and np.array_equal(er, g.events())


def test_mixed_worlds(ref_lib, oracle_lib):
    rng = np.random.default_rng(17)
    for trial in range(3):
        s_, d_, p_ = gen.mixed_world(rng, _nt(oracle_lib), with_player=(trial != 1))
        r, o = ora.World(ref_lib, s_, d_, p_), ora.World(oracle_lib, s_, d_, p_)
        for _ in range(4):
            r.step(scenes.DT_60HZ, 30, 1); o.step(scenes.DT_60HZ, 30, 1)
            for cls in (0, 1, 2):
                assert gen.state_equal(r.read(cls), o.read(cls))
                assert gen.same_bits(r.read(cls)["node_xform"], o.read(cls)["node_xform"])
            assert np.array_equal(r.events(), o.events())


def test_grid_pairs_equal_all_pairs_definition(ref_lib, oracle_lib):
    """Broadphase pair set = unordered dynamic pairs whose interval_collision is true on the frozen
    state: the grid-accelerated port must return exactly the reference's set."""
    import ctypes as C
    n, L = 2000, 20.0
    pos, vel = scenes.sphere_worlds(1, n, L, first_world=77)
    dy = scenes.sphere_bodies(pos[0], vel[0])
    st = scenes.plane_bodies(scenes.box_planes(L))
    w = ora.World(oracle_lib, st, dy)
    buf = np.zeros((20000, 2), np.int32)
    cnt_grid = oracle_lib.dll.ora_world_pairs(w.h, scenes.DT_60HZ, 1, buf.ctypes.data_as(C.c_void_p), len(buf))
    grid = buf[:cnt_grid].copy()
    cnt_all = oracle_lib.dll.ora_world_pairs(w.h, scenes.DT_60HZ, 0, buf.ctypes.data_as(C.c_void_p), len(buf))
    allp = buf[:cnt_all].copy()
    assert cnt_grid == cnt_all > 0 and np.array_equal(grid, allp)
    # spot-check the definition against the reference's public interval_collision
    core = ora.to_core(dy)
    for i, j in allp[:40]:
        assert ref_lib.k_interval(core[i:i + 1], core[j:j + 1], 0, gen.DT)[0] == 1
    rng = np.random.default_rng(1)
    for _ in range(300):
        i, j = sorted(rng.choice(n, 2, replace=False))
        expect = int(((allp[:, 0] == i) & (allp[:, 1] == j)).any())
        assert ref_lib.k_interval(core[i:i + 1], core[j:j + 1], 0, gen.DT)[0] == expect
