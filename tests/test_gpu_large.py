"""GPU parity tests of the large-world path (uniform grid + order-preserving fixed point) against the
oracle's sequential sweep: same state bits, same contact events in the same order, same pair sets."""
import ctypes as C

import numpy as np
import pytest

import gen
import ora
from crux_b200 import scenes

pytestmark = pytest.mark.gpu
DT = scenes.DT_60HZ


def _world(oracle_lib, n, L, seed):
    pos, vel = scenes.sphere_worlds(1, n, L, first_world=seed)
    st = scenes.plane_bodies(scenes.box_planes(L))
    return pos[0], vel[0], st, ora.World(oracle_lib, st, scenes.sphere_bodies(pos[0], vel[0]))


def _ev(e):
    return [(int(x["step"]), int(x["a_class"]), int(x["a_index"]), int(x["b_class"]), int(x["b_index"])) for x in e]


@pytest.mark.parametrize("n,L,steps", [(256, 10.0, 120), (3000, 16.0, 40), (20000, 40.0, 12)])
def test_large_world_equals_sequential_sweep(gpu_lib, oracle_lib, n, L, steps):
    pos, vel, st, o = _world(oracle_lib, n, L, 900 + n)
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    total_events = 0
    for s in range(steps):
        g.step(DT, 1)
        oracle_lib.dll.ora_world_step_grid(o.h, DT, 1, 0)
        ex = g.exactness()
        assert ex["proven_exact"], (s, ex)
        P, V, R = g.download_state()
        r = o.read(1)
        assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"]) and np.array_equal(R != 0, r["at_rest"] != 0), (s, ex)
        ge, oe = _ev(g.events()), _ev(o.events())
        assert ge == oe, (s, len(ge), len(oe))
        total_events += len(ge)
    assert total_events > 0
    g.close()


def test_large_world_multi_step_call(gpu_lib, oracle_lib):
    n, L = 5000, 24.0
    pos, vel, st, o = _world(oracle_lib, n, L, 31)
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    g.step(DT, 25)
    oracle_lib.dll.ora_world_step_grid(o.h, DT, 25, 0)
    P, V, R = g.download_state()
    r = o.read(1)
    assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"])
    assert _ev(g.events()) == _ev(o.events())
    s = g.stats()
    assert s["bodies"] == n and s["nonfinite"] == 0
    g.close()


@pytest.mark.parametrize("n,L", [(2000, 20.0), (30000, 48.0)])
def test_broadphase_pair_set_bit_exact(gpu_lib, oracle_lib, n, L):
    """Unordered pairs with interval_collision true on the frozen state, canonically sorted."""
    pos, vel, st, o = _world(oracle_lib, n, L, 77)
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    got = g.pairs(DT)
    buf = np.zeros((200000, 2), np.int32)
    cnt = oracle_lib.dll.ora_world_pairs(o.h, DT, 1, buf.ctypes.data_as(C.c_void_p), len(buf))
    assert cnt > 0 and np.array_equal(got, buf[:cnt])
    # and again after the world has evolved (contacts, resting bodies)
    g.step(DT, 30)
    oracle_lib.dll.ora_world_step_grid(o.h, DT, 30, 0)
    got = g.pairs(DT)
    cnt = oracle_lib.dll.ora_world_pairs(o.h, DT, 1, buf.ctypes.data_as(C.c_void_p), len(buf))
    assert np.array_equal(got, buf[:cnt])
    g.close()
