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


@pytest.mark.parametrize("n,L,steps,env", [(256, 10.0, 120, {}), (3000, 16.0, 40, {}), (20000, 40.0, 12, {}),
                                           (3000, 16.0, 40, {"PG_LARGE_ITER_BATCH": "1"}),       # one fixed-point iteration per host read-back
                                           (3000, 12.0, 40, {"PG_LARGE_TINY_BUFFERS": "1"}),     # candidate list, ball buffer and hash set all have to grow
                                           (3000, 12.0, 30, {"PG_LARGE_SMALL_STAGES": "1"}),     # overflow paths of the block stages
                                           (3000, 12.0, 30, {"PG_LARGE_PASS3_SPLIT_MIN": "0"})]) # pass 3 as probe + dense exact kernel (large worlds)
def test_large_world_equals_sequential_sweep(gpu_lib, oracle_lib, monkeypatch, n, L, steps, env):
    for k in ("PG_LARGE_ITER_BATCH", "PG_LARGE_TINY_BUFFERS", "PG_LARGE_SMALL_STAGES", "PG_LARGE_PASS3_SPLIT_MIN"):
        monkeypatch.delenv(k, raising=False)
    for k, val in env.items():
        monkeypatch.setenv(k, val)                 # read by pg_large_create
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


def test_large_world_with_stragglers_and_fast_bodies(gpu_lib, oracle_lib):
    """Bodies far outside the bulk (tunnelled through the floor and falling, or simply far away) and
    bodies much faster than the rest: the grid is sized for the bulk, the odd ones search wider or are
    checked from the fast list -- and the result is still the sequential sweep's, bit for bit."""
    n, L = 3000, 16.0
    pos, vel = scenes.sphere_worlds(1, n, L, first_world=4242)
    pos, vel = pos[0].copy(), vel[0].copy()
    rng = np.random.default_rng(9)
    fall = rng.choice(n, 40, replace=False)
    pos[fall, 1] = -rng.uniform(50.0, 900.0, 40).astype(np.float32)
    vel[fall, 1] = -rng.uniform(30.0, 170.0, 40).astype(np.float32)
    pos[fall[:6], 0] = pos[fall[6:12], 0] + np.float32(0.31)          # some of them close enough to meet on the way down
    pos[fall[:6], 1] = pos[fall[6:12], 1] + np.float32(0.05)
    pos[fall[:6], 2] = pos[fall[6:12], 2]
    far = rng.choice(np.setdiff1d(np.arange(n), fall), 5, replace=False)
    pos[far, 0] = np.float32(1.0e4) + np.arange(5, dtype=np.float32)
    fast = rng.choice(np.setdiff1d(np.arange(n), np.concatenate([fall, far])), 25, replace=False)
    vel[fast] *= np.float32(20.0)                                       # up to 80 m/s inside the box
    st = scenes.plane_bodies(scenes.box_planes(L))
    o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos, vel))
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    total = 0
    for s in range(30):
        g.step(DT, 1)
        oracle_lib.dll.ora_world_step_grid(o.h, DT, 1, 0)
        assert g.exactness()["proven_exact"], s
        P, V, R = g.download_state()
        r = o.read(1)
        assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"]) and np.array_equal(R != 0, r["at_rest"] != 0), s
        ge, oe = _ev(g.events()), _ev(o.events())
        assert ge == oe, (s, len(ge), len(oe))
        total += len(ge)
    assert total > 500
    g.close()


def test_large_world_grazing_carpet(gpu_lib, oracle_lib):
    """A carpet of touching spheres with tiny velocities: grazing pairs whose interval search runs to hundreds
    or thousands of nodes.  One thread per pair gives them up after 48 nodes; a block per pair shares the leaf
    scan (k_eval_hard) -- and the result is still the sequential sweep's, bit for bit."""
    side = 32
    n = side * side
    g1 = np.arange(side, dtype=np.float32)
    gx, gz = np.meshgrid(g1, g1)
    pos = np.stack([(gx.ravel() - 15.5) * np.float32(0.3000001), np.full(n, 0.151, np.float32) + (np.arange(n) % 3) * np.float32(1.0e-6),
                    (gz.ravel() - 15.5) * np.float32(0.3000001)], axis=1).astype(np.float32)
    rng = np.random.default_rng(5)
    vel = (rng.standard_normal((n, 3)) * np.float32(0.02)).astype(np.float32)
    vel[::5] *= np.float32(100.0)
    st = scenes.plane_bodies(scenes.box_planes(12.0))
    o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos, vel))
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    total = 0
    for s in range(24):
        g.step(DT, 1)
        oracle_lib.dll.ora_world_step_grid(o.h, DT, 1, 0)
        ex = g.exactness()
        assert ex["proven_exact"], (s, ex)
        P, V, R = g.download_state()
        r = o.read(1)
        assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"]) and np.array_equal(R != 0, r["at_rest"] != 0), (s, ex)
        ge, oe = _ev(g.events()), _ev(o.events())
        assert ge == oe, (s, len(ge), len(oe))
        total += len(ge)
    assert total > 1000, total
    g.close()


def test_large_world_random_campaign(gpu_lib, oracle_lib):
    """Random size, density, speed, per-body radii, restitution and time step: state, contact order and the
    exactness certificate, step by step, against the sequential sweep."""
    rng = np.random.default_rng(21)
    total = 0
    unproven = []
    for case in range(16):
        n = int(rng.choice([64, 300, 1000, 2500]))
        L = float(rng.choice([3.0, 6.0, 12.0, 30.0]))
        while n * 0.0654 > 0.25 * L ** 3:        # keep the packing fraction below 25 % even with the largest radius (0.25):
            L *= 2.0                             # a crush of overlapping spheres exceeds the path's documented limits
        speed = float(rng.choice([1.0, 8.0, 30.0]))
        restitution = float(rng.choice([1.0, 0.5, 0.0]))
        dt = np.float32(rng.choice([1.0 / 60.0, 0.004]))
        pos, vel = scenes.sphere_worlds(1, n, L, first_world=7000 + case, speed=speed)
        radius = rng.choice([0.05, 0.15, 0.25], size=n).astype(np.float32) if case % 2 else np.float32(0.15)
        st = scenes.plane_bodies(scenes.box_planes(L))
        dyn = scenes.sphere_bodies(pos[0], vel[0], restitution=restitution)
        dyn["shape"][:, 3] = radius
        o = ora.World(oracle_lib, st, dyn)
        g = gpu_lib.GpuLargeWorld(n, len(st))
        g.set(st, pos[0], vel[0], radius=radius, restitution=restitution)
        tag = (case, n, L, speed, restitution, float(dt))
        for s in range(12):
            try:
                g.step(dt, 1)
            except gpu_lib.PgUnproven:
                # a crush of fast spheres can need more fixed-point iterations than the path allows; the step then
                # returns PG_ERR_UNPROVEN -- loudly, never PG_OK -- and the case ends here
                assert not g.exactness()["proven_exact"]
                unproven.append(tag + (s, g.exactness()["iterations"]))
                break
            oracle_lib.dll.ora_world_step_grid(o.h, dt, 1, 0)
            ex = g.exactness()
            assert ex["proven_exact"], tag + (s,)
            P, V, R = g.download_state()
            r = o.read(1)
            assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"]) and np.array_equal(R != 0, r["at_rest"] != 0), tag + (s,)
            ge, oe = _ev(g.events()), _ev(o.events())
            assert ge == oe, tag + (s, len(ge), len(oe))
            total += len(ge)
        g.close()
    assert total > 2000 and len(unproven) == 0, unproven


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


@pytest.mark.parametrize("with_player,split", [(False, False), (True, False), (False, True)])
def test_level_geometry_world_c5_style(gpu_lib, oracle_lib, monkeypatch, with_player, split):
    """Config C5 in small: box planes + static AABB slabs (scene nodes) + spheres visited through their
    scene nodes + the player capsule against the level; state, node semantics and event order.
    split: pass 3 as probe kernel + dense exact kernel, as worlds of 262144 bodies and more run it."""
    if split:
        monkeypatch.setenv("PG_LARGE_PASS3_SPLIT_MIN", "0")
    else:
        monkeypatch.delenv("PG_LARGE_PASS3_SPLIT_MIN", raising=False)
    n, L = 2500, 32.0
    statics, pos, vel, players = scenes.level_world(n, L, seed_world=3, with_player=with_player)
    # make sure spheres are actually moving into slabs: a denser, faster subset near the floors
    vel = (vel * np.float32(1.5)).astype(np.float32)
    g = gpu_lib.GpuLargeWorld(n, len(statics))
    g.set(gpu_lib.prepare(statics), pos, vel)
    if with_player:
        pl = gpu_lib.prepare(players)
        pl["node_xform"][0] = gpu_lib.node_transform(pl["position"][0], pl["rotation"][0], pl["scale"][0], pl["parent_xform"][0])
        g.set_players(pl)
        players = pl
    o = ora.World(oracle_lib, statics, scenes.sphere_bodies_with_nodes(pos, vel), players if with_player else None)
    box_events = 0
    for s in range(25):
        g.step(DT, 1)
        o.step(DT, 1, 1)
        assert g.exactness()["proven_exact"], s
        P, V, R = g.download_state()
        r = o.read(1)
        assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"]) and np.array_equal(R != 0, r["at_rest"] != 0), s
        ge, oe = g.events(), o.events()
        key = lambda e: [(int(x["a_class"]), int(x["a_index"]), int(x["b_class"]), int(x["b_index"])) for x in e]
        assert key(ge) == key(oe), (s, len(ge), len(oe))
        box_events += int(((oe["a_class"] == 0) & (oe["b_class"] == 1)).sum())
        if with_player:
            gp, op = g.get_players(), o.read(2)
            assert gen.state_equal(gp, op), s
    assert box_events > 0          # AABB-sphere contacts really happened
    g.close()


# ------------------------------------------------------------------------------------------------
# parity at the CONFIGURED sizes (BASELINE.json configs C2, C4, C5) against the oracle's grid-accelerated
# restatement of the sequential sweep (ora_world_step_grid: same visits in the same order, proven equal to the
# all-pairs loop in tests/test_oracle_vs_ref.py; it falls back to the all-pairs loop by itself when its bounds
# are violated)
# ------------------------------------------------------------------------------------------------

def _compare_step(g, o, oracle_lib, tag):
    g.step(DT, 1)
    oracle_lib.dll.ora_world_step_grid(o.h, DT, 1, 0)
    ex = g.exactness()
    assert ex["proven_exact"], (tag, ex)
    P, V, R = g.download_state()
    r = o.read(1)
    assert gen.same_bits(P, r["position"]) and gen.same_bits(V, r["velocity"]) and np.array_equal(R != 0, r["at_rest"] != 0), (tag, ex)
    ge, oe = g.events(), o.events()
    assert len(ge) == len(oe), (tag, len(ge), len(oe))
    for f in ("a_class", "a_index", "b_class", "b_index"):
        assert np.array_equal(ge[f], oe[f]), (tag, f)
    return len(ge)


def test_c2_full_size_three_steps_and_pair_set(gpu_lib, oracle_lib):
    """Config C2 as BASELINE.json states it: one world, 65,536 spheres, L = 64.  Three steps bit for bit (state,
    at_rest, contact events in solve order) and the broadphase pair set of the state after them."""
    n, L = 65536, 64.0
    pos, vel, st, o = _world(oracle_lib, n, L, 0)
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    events = sum(_compare_step(g, o, oracle_lib, ("c2", s)) for s in range(3))
    assert events > 1000
    got = g.pairs(DT)
    buf = np.zeros((max(4 * len(got), 1024), 2), np.int32)
    cnt = oracle_lib.dll.ora_world_pairs(o.h, DT, 1, buf.ctypes.data_as(C.c_void_p), len(buf))
    assert cnt == len(got) and np.array_equal(got, buf[:cnt])
    g.close()


def test_c2_full_size_from_a_thermalised_state(gpu_lib, oracle_lib):
    """C2 after 300 steps on the device: spheres have fallen up to 44 m, bounced back and collide at tens of m/s
    (the regime in which the previous candidate test listed hundreds of neighbours per body and the fixed point
    needed 22 iterations).  The oracle is started from the downloaded state; two more steps, bit for bit."""
    n, L = 65536, 64.0
    pos, vel = scenes.sphere_worlds(1, n, L, first_world=0)
    st = scenes.plane_bodies(scenes.box_planes(L))
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos[0], vel[0])
    g.step(DT, 300)
    P, V, R = g.download_state()
    assert np.isfinite(P).all() and float(np.linalg.norm(V, axis=1).max()) > 20.0
    dyn = scenes.sphere_bodies(P, V)
    dyn["at_rest"] = R
    o = ora.World(oracle_lib, st, dyn)
    events = sum(_compare_step(g, o, oracle_lib, ("c2-thermal", s)) for s in range(2))
    assert events > 1000
    g.close()


def test_c4_full_size_one_step(gpu_lib, oracle_lib):
    """Config C4: one world, 4,194,304 spheres, L = 256: one step, every body and every contact event."""
    n, L = 4194304, 256.0
    pos, vel, st, o = _world(oracle_lib, n, L, 0)
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    events = _compare_step(g, o, oracle_lib, "c4")
    assert events > 20000
    g.close()


def test_c5_full_size_one_step_regions(gpu_lib, oracle_lib):
    """Config C5: 1,048,576 spheres + the slab level (static AABBs with scene nodes) + the player capsule,
    one step on the device at full size.  The CPU sweep over 1 M spheres x thousands of slabs takes hours, so
    the check is regional: all spheres (array order kept) and all statics (array order kept) that reach into a
    box are stepped by the oracle's all-pairs sweep, and the spheres well inside it -- farther from its faces
    than any chain of contacts can carry in one step -- must come out with the same bits.  Pass 3 of EVERY
    sphere is independent of the others and is covered by a second, random sample."""
    n, L = 1048576, 160.0
    statics, pos, vel, players = scenes.level_world(n, L, seed_world=0, with_player=True)
    g = gpu_lib.GpuLargeWorld(n, len(statics))
    g.set(gpu_lib.prepare(statics), pos, vel)
    pl = gpu_lib.prepare(players)
    pl["node_xform"][0] = gpu_lib.node_transform(pl["position"][0], pl["rotation"][0], pl["scale"][0], pl["parent_xform"][0])
    g.set_players(pl)
    g.step(DT, 1)
    assert g.exactness()["proven_exact"]
    P, V, R = g.download_state()
    ev = g.events()
    checked = 0
    for centre, half, margin in (((0.0, 4.0, 0.0), 9.0, 3.0), ((40.0, 60.0, -30.0), 9.0, 3.0)):
        c = np.array(centre, np.float32)
        inside = np.all(np.abs(pos - c) <= half, axis=1)
        deep = np.all(np.abs(pos - c) <= half - margin, axis=1)
        idx = np.nonzero(inside)[0]
        # statics: the six planes, and the slabs whose box comes within reach of the region
        sc = statics["node_xform"][:, 12:15] + statics["shape"][:, 0:3]
        se = statics["shape"][:, 3:6]
        near = np.all(np.abs(sc - c) <= se + half + 1.0, axis=1) | (statics["collider_type"] == scenes.COLLIDER_PLANE)
        sidx = np.nonzero(near)[0]
        o = ora.World(oracle_lib, statics[sidx], scenes.sphere_bodies_with_nodes(pos[idx], vel[idx]), None)
        o.step(DT, 1, 1)
        r = o.read(1)
        d = deep[idx]
        assert d.sum() > 200
        assert gen.same_bits(P[idx][d], r["position"][d]) and gen.same_bits(V[idx][d], r["velocity"][d]) and \
            np.array_equal(R[idx][d] != 0, r["at_rest"][d] != 0), centre
        # contact events whose spheres are all deep: the same sequence (array order is kept inside the region, so
        # the region's solve order restricted to those bodies is the world's)
        deep_global = set(idx[d].tolist())

        def keep(cls_a, ia, cls_b, ib):
            spheres = [i for c_, i in ((cls_a, ia), (cls_b, ib)) if c_ == 1]
            return len(spheres) > 0 and all(i in deep_global for i in spheres)
        want = []
        for e in o.events():
            ia = int(sidx[e["a_index"]]) if e["a_class"] == 0 else int(idx[e["a_index"]])
            ib = int(sidx[e["b_index"]]) if e["b_class"] == 0 else int(idx[e["b_index"]])
            if keep(int(e["a_class"]), ia, int(e["b_class"]), ib):
                want.append((int(e["a_class"]), ia, int(e["b_class"]), ib))
        got = [(int(e["a_class"]), int(e["a_index"]), int(e["b_class"]), int(e["b_index"])) for e in ev
               if e["a_class"] != 2 and e["b_class"] != 2 and keep(int(e["a_class"]), int(e["a_index"]), int(e["b_class"]), int(e["b_index"]))]
        assert got == want, (centre, len(got), len(want))
        checked += int(d.sum())
    assert checked > 500
    # the player capsule against the level (passes 1-2): the any-collider kernel's result against the oracle's
    o = ora.World(oracle_lib, statics, None, pl)
    o.step(DT, 1, 1)
    assert gen.state_equal(g.get_players(), o.read(2))
    g.close()


def test_tall_column_through_free_fall_compression_and_hot_gas(gpu_lib, oracle_lib):
    """The regimes the configured C4 run goes through (spheres fall for seconds, the first ones bounce back into
    the ones still falling at a relative 100 m/s, then a hot gas), in a column small enough for the oracle: a
    12 x 256 x 12 box at C4's number density (9216 spheres).  The device advances the world; at six points of
    the run -- free fall, compression front, hot gas with bodies that have tunnelled through the floor -- the
    oracle is started from the downloaded state and three steps are compared bit for bit (state, at_rest,
    contact events in solve order).  Every step of the run must be proven exact."""
    W, H = 12.0, 256.0
    n = int(W * W * H * 0.25)
    rng = np.random.default_rng(404)
    pos = np.stack([rng.uniform(-W / 2 + 0.5, W / 2 - 0.5, n), rng.uniform(0.5, H - 0.5, n), rng.uniform(-W / 2 + 0.5, W / 2 - 0.5, n)], axis=1).astype(np.float32)
    vel = rng.uniform(-4.0, 4.0, (n, 3)).astype(np.float32)
    planes = np.array([[0, 1, 0, 0], [0, -1, 0, -H], [1, 0, 0, -W / 2], [-1, 0, 0, -W / 2], [0, 0, 1, -W / 2], [0, 0, -1, -W / 2]], np.float32)
    st = scenes.plane_bodies(planes)
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    done, total_events, max_iter = 0, 0, 0
    for target in (60, 200, 300, 360, 450, 700):
        g.step(DT, target - done)          # raises PgUnproven if any step's fixed point is not proven
        done = target
        P, V, R = g.download_state()
        assert np.isfinite(P).all() and np.isfinite(V).all()
        dyn = scenes.sphere_bodies(P, V)
        dyn["at_rest"] = R
        o = ora.World(oracle_lib, st, dyn)
        for s in range(3):
            total_events += _compare_step(g, o, oracle_lib, ("column", target, s))
            max_iter = max(max_iter, g.exactness()["iterations"])
        done += 3
        o.close()
    assert total_events > 3000 and max_iter >= 6, (total_events, max_iter)
    assert float(np.linalg.norm(V, axis=1).max()) > 40.0 and int((P[:, 1] < -1.0).sum()) > 0      # hot, and some tunnelled out
    g.close()


def test_failed_step_leaves_the_state_untouched_and_bad_input_is_refused(gpu_lib, oracle_lib):
    """(i) A step that cannot be completed -- here: one sphere in the middle of 40 that overlap it, more contacts than a
    timeline holds -- returns PG_ERR_CAPACITY and puts the state back bit for bit (pass 3 had already written it);
    the handle stays usable.  (ii) Non-capsule players are refused (a sphere player has live table entries against
    the dynamic spheres which the large-world path does not evaluate).  (iii) The binding refuses arrays of the wrong
    dtype / shape / layout instead of letting the library copy past them."""
    n, L = 600, 12.0
    pos, vel = scenes.sphere_worlds(1, n, L, first_world=55)
    pos, vel = pos[0].copy(), vel[0].copy()
    rng = np.random.default_rng(8)
    pos[:41] = np.array([0.0, 5.0, 0.0], np.float32) + rng.normal(0, 0.03, (41, 3)).astype(np.float32)     # a crush of 41 overlapping spheres
    vel[:41] = rng.normal(0, 1.0, (41, 3)).astype(np.float32)
    st = scenes.plane_bodies(scenes.box_planes(L))
    g = gpu_lib.GpuLargeWorld(n, len(st))
    g.set(st, pos, vel)
    with pytest.raises(gpu_lib.PgError, match=r"\(-3\)"):
        g.step(DT, 1)
    P, V, R = g.download_state()
    assert gen.same_bits(P, pos) and gen.same_bits(V, vel) and not R.any()
    assert len(g.events()) == 0
    # the same handle, a sane world: steps and matches the oracle
    pos2, vel2 = scenes.sphere_worlds(1, n, L, first_world=56)
    g.set(st, pos2[0], vel2[0])
    o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos2[0], vel2[0]))
    for s in range(5):
        _compare_step(g, o, oracle_lib, ("after failure", s))
    # (iii)
    with pytest.raises(gpu_lib.PgError):
        g.upload_state(pos2[0].astype(np.float64), vel2[0])
    with pytest.raises(gpu_lib.PgError):
        g.upload_state(pos2[0][:-1], vel2[0][:-1])
    with pytest.raises(gpu_lib.PgError):
        g.download_state(np.zeros((n, 3), np.float32)[:, ::-1], np.zeros((n, 3), np.float32), np.zeros(n, np.uint8))
    g.close()
    # (ii)
    statics, p5, v5, players = scenes.level_world(500, 32.0, seed_world=1, with_player=True)
    g = gpu_lib.GpuLargeWorld(500, len(statics))
    g.set(gpu_lib.prepare(statics), p5, v5)
    bad = gpu_lib.prepare(players.copy())
    bad["collider_type"] = scenes.COLLIDER_SPHERE
    with pytest.raises(gpu_lib.PgError, match=r"\(-4\)"):
        g.set_players(bad)
    g.set_players(gpu_lib.prepare(players))
    g.close()
