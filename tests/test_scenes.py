import numpy as np

from crux_b200 import scenes


def test_lcg_matches_scalar_definition():
    """s = s*1664525 + 1013904223 (mod 2^32), u = (s >> 8) / 2^24 -- SURVEY.md section 8(d)."""
    seeds = [12345, 12346, 99999]
    u = scenes.lcg_uniform(np.array(seeds), 50)
    for w, seed in enumerate(seeds):
        s = seed
        for k in range(50):
            s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
            assert u[w, k] == np.float32((s >> 8) / 16777216.0)


def test_sphere_worlds_are_shardable():
    """A rank that owns worlds [a, b) generates exactly the slice of the global batch."""
    full_p, full_v = scenes.sphere_worlds(8, 16, 10.0)
    part_p, part_v = scenes.sphere_worlds(3, 16, 10.0, first_world=4)
    assert np.array_equal(full_p[4:7], part_p) and np.array_equal(full_v[4:7], part_v)
    assert full_p[..., 1].min() >= 0.5 and np.abs(full_p[..., 0]).max() <= 4.5 and np.abs(full_v).max() <= 4.0


def test_box_is_closed_and_normals_point_inwards():
    pl = scenes.box_planes(10.0)
    centre = np.array([0, 5, 0], np.float32)
    assert np.all(pl[:, :3] @ centre - pl[:, 3] > 0)          # centre on the positive side of all six
    assert np.allclose(np.abs(pl[:, :3]).sum(axis=1), 1.0)
