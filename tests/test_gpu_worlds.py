"""GPU parity tests: the CUDA path against the oracle, through the C-ABI (tests marked gpu)."""
import numpy as np
import pytest

import gen
import ora
from crux_b200 import scenes

pytestmark = pytest.mark.gpu
DT = scenes.DT_60HZ


def test_pair_entry_points_bit_exact(gpu_lib, oracle_lib):
    """All 7 live collider pairs: distance, interval gate, narrowphase, resolver, full visit."""
    rng = np.random.default_rng(11)
    n = 2100
    A, B = gen.rand_pairs(rng, n, gpu_lib.node_transform)
    t = rng.uniform(0, gen.DT, n).astype(np.float32)
    d_gpu = gpu_lib.pair_min_dist(A, B, t)
    hit_gpu, _ = gpu_lib.pair_interval_collision(A, B, 0.0, gen.DT)
    col_gpu, ht_gpu = gpu_lib.pair_narrow_phase(A, B, gen.DT)
    ev_gpu, Av, Bv = gpu_lib.pair_visit(A, B, gen.DT)
    mm_gpu = gpu_lib.pair_max_move(A, 0.0, gen.DT)
    n_hits = 0
    for k in range(n):
        a, b = A[k:k + 1], B[k:k + 1]
        assert gen.same_bits(d_gpu[k], oracle_lib.k_min_dist(a, b, t[k])), k
        assert gen.same_bits(mm_gpu[k], oracle_lib.k_max_move(a, 0.0, gen.DT)), k
        assert hit_gpu[k] == oracle_lib.k_interval(a, b, 0.0, gen.DT)[0], k
        c, ht = oracle_lib.k_narrow(a, b, gen.DT)
        assert col_gpu[k] == c and gen.same_bits(ht_gpu[k], ht), k
        ev, ao, bo = oracle_lib.k_visit(a, b, gen.DT)
        assert ev_gpu[k] == ev, k
        assert gen.state_equal(Av[k:k + 1], ao) and gen.state_equal(Bv[k:k + 1], bo), k
        n_hits += ev
    assert n_hits > 300          # the generator really exercises the resolvers
    # resolver alone, at the narrowphase time where it is valid, else a fixed time
    ht = np.where(ht_gpu >= 0, ht_gpu, np.float32(0.004)).astype(np.float32)
    Ar, Br = gpu_lib.pair_resolve(A, B, ht, gen.DT)
    for k in range(0, n, 3):
        _, ao, bo = oracle_lib.k_resolve(A[k:k + 1], B[k:k + 1], ht[k], gen.DT)
        assert gen.state_equal(Ar[k:k + 1], ao) and gen.state_equal(Br[k:k + 1], bo), k


def _events_tuple(ev):
    return [(int(e["step"]), int(e["a_class"]), int(e["a_index"]), int(e["b_class"]), int(e["b_index"]), int(e["event_type"])) for e in ev]


def test_general_kernel_sphere_world_matches_oracle(gpu_lib, oracle_lib):
    """256-sphere bouncehouse world through the any-collider kernel: 120 steps, state + contact order."""
    st, dy = scenes.bouncehouse_world(3)
    w = gpu_lib.GpuWorlds(1, len(st), len(dy), 0, max_events_per_world=8192)
    w.set_bodies(0, scenes.CLASS_STATIC, st)
    w.set_bodies(0, scenes.CLASS_DYNAMIC, dy)
    o = ora.World(oracle_lib, st, dy)
    for chunk in range(4):
        w.step(DT, 30)
        o.step(DT, 30)
        assert gen.state_equal(w.get_bodies(0, scenes.CLASS_DYNAMIC), o.read(1)), chunk
        assert _events_tuple(w.events()) == _events_tuple(o.events()), chunk
    w.close()


def test_general_kernel_mixed_world_matches_oracle(gpu_lib, oracle_lib):
    """Planes + static AABBs + dynamic spheres/AABBs/items + a capsule player, scene nodes refreshed
    every step (the headless scene_update recipe): all four passes and every live table entry."""
    rng = np.random.default_rng(5)
    total_events = 0
    for trial in range(3):
        st, dy, pl = gen.mixed_world(rng, gpu_lib.node_transform, with_player=(trial != 1))
        w = gpu_lib.GpuWorlds(1, len(st), len(dy), len(pl), max_events_per_world=4096)
        w.set_bodies(0, scenes.CLASS_STATIC, st)
        w.set_bodies(0, scenes.CLASS_DYNAMIC, dy)
        if len(pl):
            w.set_bodies(0, scenes.CLASS_PLAYER, pl)
        o = ora.World(oracle_lib, st, dy, pl)
        for chunk in range(6):
            w.step(DT, 25, refresh_nodes=True)
            o.step(DT, 25, refresh_nodes=1)
            for cls in (0, 1, 2):
                g, r = w.get_bodies(0, cls), o.read(cls)
                assert gen.state_equal(g, r), (trial, chunk, cls)
                assert gen.same_bits(g["node_xform"], r["node_xform"]), (trial, chunk, cls)
            eg, eo = _events_tuple(w.events()), _events_tuple(o.events())
            assert eg == eo, (trial, chunk)
            total_events += len(eg)
        w.close()
    assert total_events > 50


def test_sphere_kernel_matches_oracle(gpu_lib, oracle_lib):
    """Warp-per-world kernel, 6 worlds x 256 spheres, 200 steps in chunks: bit-exact state, and the
    contact sequence of every chunk equals the oracle's, in order."""
    n_worlds, n = 6, 256
    pos, vel = scenes.sphere_worlds(n_worlds, n, 10.0, first_world=100)
    st = scenes.plane_bodies(scenes.box_planes(10.0))
    w = gpu_lib.GpuWorlds(n_worlds, len(st), n, 0, max_events_per_world=8192)
    w.set_spheres(0, st, pos, vel)
    worlds = [ora.World(oracle_lib, st, scenes.sphere_bodies(pos[k], vel[k])) for k in range(n_worlds)]
    for chunk in range(5):
        w.step(DT, 40)
        ev = w.events()
        for k, o in enumerate(worlds):
            o.step(DT, 40)
            assert gen.state_equal(w.get_bodies(k, scenes.CLASS_DYNAMIC), o.read(1)), (chunk, k)
            assert _events_tuple(ev[ev["world"] == k]) == _events_tuple(o.events()), (chunk, k)
    # packed state exchange round trip
    P = np.zeros((n_worlds, n, 3), np.float32); V = np.zeros_like(P); R = np.zeros((n_worlds, n), np.uint8)
    w.download_state(0, n_worlds, P, V, R)
    for k, o in enumerate(worlds):
        r = o.read(1)
        assert gen.same_bits(P[k], r["position"]) and gen.same_bits(V[k], r["velocity"])
        assert np.array_equal(R[k] != 0, r["at_rest"] != 0)
    w.upload_state(0, n_worlds, P, V, R)
    w.step(DT, 10)
    for k, o in enumerate(worlds):
        o.step(DT, 10)
        assert gen.state_equal(w.get_bodies(k, scenes.CLASS_DYNAMIC), o.read(1)), k
    st_ = w.stats()
    assert st_["bodies"] == n_worlds * n and st_["nonfinite"] == 0
    w.close()


@pytest.mark.parametrize("n", [1, 31, 33, 100, 500])
def test_sphere_kernel_ragged_sizes(gpu_lib, oracle_lib, n):
    """World sizes around the lane/slot boundaries of the warp kernel (and one past 256)."""
    L = 6.0 if n <= 100 else 12.0
    pos, vel = scenes.sphere_worlds(2, n, L, first_world=7)
    st = scenes.plane_bodies(scenes.box_planes(L))
    w = gpu_lib.GpuWorlds(2, len(st), n, 0, max_events_per_world=16384)
    w.set_spheres(0, st, pos, vel)
    w.step(DT, 60)
    ev = w.events()
    for k in range(2):
        o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos[k], vel[k]))
        o.step(DT, 60)
        assert gen.state_equal(w.get_bodies(k, scenes.CLASS_DYNAMIC), o.read(1)), k
        assert _events_tuple(ev[ev["world"] == k]) == _events_tuple(o.events()), k
    w.close()


def _hard_worlds():
    """Worlds that drive the sphere kernel's rare paths: a dense pile (hundreds of candidate bits per
    row block, several contacts per row, dirty rows), far-away stragglers (filter centre / per-pair
    slack), a carpet of touching spheres with tiny velocities (grazing pairs: walk budget, leaf
    scan, monotone plane decisions), and an astronomically large coordinate (the `wild` path)."""
    n = 256
    out = {}
    pos, vel = scenes.sphere_worlds(1, n, 2.6, first_world=501)
    out["dense"] = (2.6, pos[0].copy(), vel[0].copy())
    pos, vel = scenes.sphere_worlds(1, n, 1.7, first_world=504)     # spheres overlap by the dozen: > 8 contacts in one row
    out["crush"] = (1.7, pos[0].copy(), vel[0].copy())
    pos, vel = scenes.sphere_worlds(1, n, 10.0, first_world=502)
    p, v = pos[0].copy(), vel[0].copy()
    p[7] = (0.5, -1500.0, 0.25); v[7] = (0.0, -170.0, 0.0)
    p[100] = (3.0e4, 5.0, -2.0); v[100] = (40.0, 0.0, 0.0)
    p[255] = (-0.75, 2.0e5, 1.0); v[255] = (0.0, 900.0, 0.0)
    out["stragglers"] = (10.0, p, v)
    g = np.arange(16, dtype=np.float32)
    gx, gz = np.meshgrid(g, g)
    p = np.stack([(gx.ravel() - 7.5) * np.float32(0.3000001), np.full(n, 0.151, np.float32) + (np.arange(n) % 3) * np.float32(1.0e-6),
                  (gz.ravel() - 7.5) * np.float32(0.3000001)], axis=1).astype(np.float32)
    rng = np.random.default_rng(5)
    v = (rng.standard_normal((n, 3)) * np.float32(0.02)).astype(np.float32)
    v[::5] *= np.float32(100.0)
    out["carpet"] = (10.0, p, v)
    pos, vel = scenes.sphere_worlds(1, n, 10.0, first_world=503)
    p, v = pos[0].copy(), vel[0].copy()
    p[33] = (1.0e18, 3.0, 0.0)
    out["wild"] = (10.0, p, v)
    return out


@pytest.mark.parametrize("name", ["dense", "crush", "stragglers", "carpet", "wild"])
def test_sphere_kernel_hard_worlds(gpu_lib, oracle_lib, name):
    L, p, v = _hard_worlds()[name]
    st = scenes.plane_bodies(scenes.box_planes(L))
    w = gpu_lib.GpuWorlds(1, len(st), len(p), 0, max_events_per_world=65536)
    w.set_spheres(0, st, p[None], v[None])
    o = ora.World(oracle_lib, st, scenes.sphere_bodies(p, v))
    total = 0
    for chunk in range(4):
        w.step(DT, 15)
        o.step(DT, 15)
        assert gen.state_equal(w.get_bodies(0, scenes.CLASS_DYNAMIC), o.read(1)), (name, chunk)
        ge, oe = _events_tuple(w.events()), _events_tuple(o.events())
        assert ge == oe, (name, chunk, len(ge), len(oe))
        total += len(ge)
    assert total > (500 if name in ("dense", "crush", "carpet") else 20), (name, total)
    w.close()


def _random_world_campaign(gpu_lib, oracle_lib, n_cases, seed, steps):
    """Random size, box, speed, radius, restitution and time step per case; bit-exact state and contact
    order against the oracle.  Worlds of one capacity class share a handle (one kernel instance)."""
    rng = np.random.default_rng(seed)
    total = 0
    for case in range(n_cases):
        n = int(rng.choice([2, 5, 17, 32, 33, 64, 65, 100, 128, 129, 200, 256]))
        radius = float(rng.choice([0.05, 0.15, 0.3]))
        L = float(rng.uniform(max(1.5, 12 * radius), 12.0))
        speed = float(rng.choice([1.0, 8.0, 40.0]))
        restitution = float(rng.choice([1.0, 0.8, 0.3, 0.0]))
        dt = np.float32(rng.choice([1.0 / 60.0, 0.004, 1.0e-4]))
        pos, vel = scenes.sphere_worlds(1, n, L, first_world=10_000 * seed + case, radius=radius, speed=speed)
        st = scenes.plane_bodies(scenes.box_planes(L))
        w = gpu_lib.GpuWorlds(1, len(st), n, 0, max_events_per_world=65536)
        w.set_spheres(0, st, pos, vel, radius=radius, restitution=restitution)
        dyn = scenes.sphere_bodies(pos[0], vel[0], radius=radius, restitution=restitution)
        o = ora.World(oracle_lib, st, dyn)
        w.step(dt, steps)
        o.step(dt, steps)
        tag = (case, n, L, speed, radius, restitution, float(dt))
        assert gen.state_equal(w.get_bodies(0, scenes.CLASS_DYNAMIC), o.read(1)), tag
        ge, oe = _events_tuple(w.events()), _events_tuple(o.events())
        if len(oe) > 65536:                      # more contacts than the event buffer holds: the first 65536 are kept, in order
            assert w.events_overflowed and ge == oe[:65536], tag + (len(ge), len(oe))
        else:
            assert ge == oe, tag + (len(ge), len(oe))
        total += len(ge)
        w.close()
    return total


def test_sphere_kernel_random_campaign(gpu_lib, oracle_lib):
    assert _random_world_campaign(gpu_lib, oracle_lib, n_cases=40, seed=1, steps=60) > 1000


def test_sphere_kernel_tiny_dt(gpu_lib, oracle_lib):
    """dt below 1e-9: the closest-approach gates are switched off, the exact path alone decides."""
    pos, vel = scenes.sphere_worlds(1, 256, 4.0, first_world=77)
    st = scenes.plane_bodies(scenes.box_planes(4.0))
    w = gpu_lib.GpuWorlds(1, len(st), 256, 0, max_events_per_world=16384)
    w.set_spheres(0, st, pos, vel)
    o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos[0], vel[0]))
    for dt in (np.float32(5.0e-10), DT, np.float32(5.0e-10)):
        w.step(dt, 5)
        o.step(dt, 5)
        assert gen.state_equal(w.get_bodies(0, scenes.CLASS_DYNAMIC), o.read(1)), float(dt)
        assert _events_tuple(w.events()) == _events_tuple(o.events()), float(dt)
    w.close()


def test_no_gpu_fallback_is_an_error(gpu_lib):
    """Capacity and mode errors surface as exceptions, never as silent CPU work."""
    w = gpu_lib.GpuWorlds(1, 2, 4, 0)
    with pytest.raises(gpu_lib.PgError):
        w.set_bodies(0, scenes.CLASS_DYNAMIC, scenes.sphere_bodies(np.zeros((9, 3)), np.zeros((9, 3))))
    w.close()


def test_cuda_path_against_reference_fixtures(gpu_lib):
    """The CUDA path against committed outputs of the unmodified reference (tests/golden/), so the
    chain reference -> fixtures -> GPU does not depend on the port being right."""
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(G, "pairs_kat.npz"))
    A = ora.from_core(z["A"]); B = ora.from_core(z["B"])
    d = gpu_lib.pair_min_dist(A, B, z["t"])
    assert gen.same_bits(d, z["dist"])
    hit, _ = gpu_lib.pair_interval_collision(A, B, 0.0, gen.DT)
    assert np.array_equal(hit, z["interval"])
    col, ht = gpu_lib.pair_narrow_phase(A, B, gen.DT)
    assert np.array_equal(col, z["colliding"]) and gen.same_bits(ht, z["hit_time"])
    ev, Av, Bv = gpu_lib.pair_visit(A, B, gen.DT)
    assert np.array_equal(ev, z["visit"]) and gen.state_equal(Av, z["A_after"]) and gen.state_equal(Bv, z["B_after"])
    Ar, Br = gpu_lib.pair_resolve(A, B, z["resolve_time"], gen.DT)
    assert gen.state_equal(Ar, z["A_resolved"]) and gen.state_equal(Br, z["B_resolved"])

    z = np.load(os.path.join(G, "spheres_256.npz"))
    pos, vel = scenes.sphere_worlds(1, 256, 10.0)
    st = scenes.plane_bodies(scenes.box_planes(10.0))
    w = gpu_lib.GpuWorlds(1, len(st), 256, 0, max_events_per_world=4096)
    w.set_spheres(0, st, pos, vel)
    done = 0
    for target in (1, 10, 100, 1000):
        w.step(DT, target - done)
        if target == 100:
            pass
        done = target
        assert gen.state_equal(w.get_bodies(0, scenes.CLASS_DYNAMIC), z[f"state_{target}"]), target
    w.close()
    w = gpu_lib.GpuWorlds(1, len(st), 256, 0, max_events_per_world=4096)
    w.set_spheres(0, st, pos, vel)
    w.step(DT, 100)
    assert _events_tuple(w.events()) == _events_tuple(z["events_first100"])
    w.close()

    z = np.load(os.path.join(G, "mixed_world.npz"))
    g = gpu_lib.GpuWorlds(1, len(z["statics"]), len(z["dynamics"]), len(z["players"]), max_events_per_world=4096)
    g.set_bodies(0, scenes.CLASS_STATIC, ora.from_core(z["statics"]))
    g.set_bodies(0, scenes.CLASS_DYNAMIC, ora.from_core(z["dynamics"]))
    g.set_bodies(0, scenes.CLASS_PLAYER, ora.from_core(z["players"]))
    done = 0
    events = []
    for target in (25, 50, 150):
        g.step(DT, target - done, refresh_nodes=True)
        ev = g.events()
        ev["step"] += done
        events.extend(_events_tuple(ev))
        done = target
        for cls, name in ((0, "S"), (1, "D"), (2, "P")):
            assert gen.state_equal(g.get_bodies(0, cls), z[f"{name}_{target}"]), (target, name)
    assert events == _events_tuple(z["events"])
    g.close()


def test_large_batch_properties(gpu_lib, oracle_lib):
    """Full C3 size (4096 x 256): size-independent properties -- the batch result equals running
    any sampled world alone (worlds are independent), no NaNs, every body stays accounted for --
    plus the oracle on a handful of sampled worlds."""
    nw, n = 4096, 256
    pos, vel = scenes.sphere_worlds(nw, n, 10.0)
    st = scenes.plane_bodies(scenes.box_planes(10.0))
    w = gpu_lib.GpuWorlds(nw, len(st), n, 0, max_events_per_world=0)
    w.set_spheres(0, st, pos, vel)
    w.step(DT, 50)
    P = np.zeros((nw, n, 3), np.float32); V = np.zeros_like(P); R = np.zeros((nw, n), np.uint8)
    w.download_state(0, nw, P, V, R)
    s = w.stats()
    assert s["bodies"] == nw * n and s["nonfinite"] == 0 and np.isfinite(P).all()
    from crux_b200 import shard
    want = shard.stats_from_state(V, R)
    assert np.isclose(s["kinetic"], want[1], rtol=1e-9) and s["at_rest"] == want[6]
    for k in (0, 1, 2047, 4095):
        o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos[k], vel[k]))
        o.step(DT, 50)
        r = o.read(1)
        assert gen.same_bits(P[k], r["position"]) and gen.same_bits(V[k], r["velocity"]), k
    w.close()


def test_step_host_pipeline_matches_resident_path(gpu_lib, oracle_lib):
    """pg_worlds_step_host (chunked H2D / step / D2H pipeline) gives the same bits as stepping the
    resident state, and as the oracle on sampled worlds."""
    nw, n = 37, 256           # not a multiple of the chunk count
    pos, vel = scenes.sphere_worlds(nw, n, 10.0, first_world=500)
    st = scenes.plane_bodies(scenes.box_planes(10.0))
    a = gpu_lib.GpuWorlds(nw, len(st), n, 0); a.set_spheres(0, st, pos, vel)
    b = gpu_lib.GpuWorlds(nw, len(st), n, 0); b.set_spheres(0, st, pos, vel)
    P = pos.copy(); V = vel.copy(); R = np.zeros((nw, n), np.uint8)
    for _ in range(6):
        a.step_host(DT, 5, P, V, R)
    b.step(DT, 30)
    P2 = np.zeros_like(P); V2 = np.zeros_like(V); R2 = np.zeros_like(R)
    b.download_state(0, nw, P2, V2, R2)
    assert gen.same_bits(P, P2) and gen.same_bits(V, V2) and np.array_equal(R, R2)
    for k in (0, 17, 36):
        o = ora.World(oracle_lib, st, scenes.sphere_bodies(pos[k], vel[k]))
        o.step(DT, 30)
        r = o.read(1)
        assert gen.same_bits(P[k], r["position"]) and gen.same_bits(V[k], r["velocity"])
    # the packed host layout (one record per world, one DMA per chunk and direction): same bits again, with a
    # ragged sphere count (n = 100: the record is padded to 16 bytes) and n_steps = 0 as the identity
    for nn in (256, 100):
        pos2, vel2 = scenes.sphere_worlds(nw, nn, 10.0, first_world=500)
        c = gpu_lib.GpuWorlds(nw, len(st), nn, 0); c.set_spheres(0, st, pos2, vel2)
        d = gpu_lib.GpuWorlds(nw, len(st), nn, 0); d.set_spheres(0, st, pos2, vel2)
        rec = c.packed_dtype()
        assert rec.itemsize % 16 == 0 and rec.itemsize >= 25 * nn
        S = np.zeros(nw, rec)
        S["pos"] = pos2; S["vel"] = vel2
        c.step_host_packed(DT, 0, S)
        assert gen.same_bits(S["pos"], pos2) and gen.same_bits(S["vel"], vel2) and not S["at_rest"].any()
        for _ in range(6):
            c.step_host_packed(DT, 5, S)
        d.step(DT, 30)
        P3 = np.zeros_like(pos2); V3 = np.zeros_like(vel2); R3 = np.zeros((nw, nn), np.uint8)
        d.download_state(0, nw, P3, V3, R3)
        assert gen.same_bits(S["pos"], P3) and gen.same_bits(S["vel"], V3) and np.array_equal(S["at_rest"], R3)
        if nn == 256:
            assert gen.same_bits(S["pos"], P)
        with pytest.raises(gpu_lib.PgError):
            c.step_host_packed(DT, 1, S[:-1])
        c.close(); d.close()
    a.close(); b.close()


def test_remove_add_and_poke_between_steps(gpu_lib, oracle_lib):
    """pg_worlds_remove_body (world.c:147-172 swap-and-pop, changes the solve order), pg_worlds_add_body
    (world.c:41-142 add-time rules) and the dirty-range poke of the player record (player.c:160-173:
    at_rest = false, velocity[1] = 3, plus a position write) between step calls, against the oracle's
    world_remove / world_add_many / world_poke: state, node matrices and event order stay bit-exact."""
    rng = np.random.default_rng(11)
    st, dy, pl = gen.mixed_world(rng, gpu_lib.node_transform, with_player=True, n_dyn=20)
    w = gpu_lib.GpuWorlds(1, len(st) + 2, len(dy) + 2, len(pl), max_events_per_world=4096)
    w.set_bodies(0, scenes.CLASS_STATIC, st)
    w.set_bodies(0, scenes.CLASS_DYNAMIC, dy)
    w.set_bodies(0, scenes.CLASS_PLAYER, pl)
    o = ora.World(oracle_lib, st, dy, pl)
    total = 0

    def run_and_compare(tag, frames=30):
        nonlocal total
        w.step(DT, frames, refresh_nodes=True)
        o.step(DT, frames, refresh_nodes=1)
        for cls in (0, 1, 2):
            g, r = w.get_bodies(0, cls), o.read(cls)
            assert len(g) == len(r), (tag, cls, len(g), len(r))
            assert gen.state_equal(g, r), (tag, cls)
            assert gen.same_bits(g["node_xform"], r["node_xform"]), (tag, cls)
        eg, eo = _events_tuple(w.events()), _events_tuple(o.events())
        assert eg == eo, tag
        total += len(eg)

    run_and_compare("start")
    # remove from the middle (the last record moves into the hole), then the last, then the first
    for step, index in enumerate((3, len(dy) - 2, 0)):
        w.remove_body(0, scenes.CLASS_DYNAMIC, index)
        o.remove(scenes.CLASS_DYNAMIC, index)
        assert w.body_count(0, scenes.CLASS_DYNAMIC) == o.counts()[1]
        run_and_compare(("remove", step))
    # a static box disappears as well (pass 1 and pass 3 lose a column)
    w.remove_body(0, scenes.CLASS_STATIC, 7)
    o.remove(scenes.CLASS_STATIC, 7)
    run_and_compare("remove static")
    # physics_add_body: a new dynamic AABB (its velocity kept) and a new static one (velocity forced to 0)
    nb = dy[5:6].copy()
    nb["position"] = [[0.3, 2.5, -0.4]]; nb["velocity"] = [[1.5, -2.0, 0.5]]
    nb["node_xform"] = gen.node_xform(gpu_lib.node_transform, nb["position"][0], [0, 0, 0], [1, 1, 1], np.eye(4, dtype=np.float32).reshape(16))
    ns = st[7:8].copy()
    ns["velocity"] = [[9.0, 9.0, 9.0]]
    assert w.add_body(0, nb[0]) == o.counts()[1]
    o.add(nb)
    assert w.add_body(0, ns[0]) == o.counts()[0]
    o.add(ns)
    assert np.all(w.get_bodies(0, scenes.CLASS_STATIC)["velocity"][-1] == 0.0)
    run_and_compare("add")
    # the player write path: jump + walk, several times
    for k in range(4):
        rec = w.get_bodies(0, scenes.CLASS_PLAYER)[0:1].copy()
        rec["at_rest"] = 0
        rec["velocity"][0, 1] = 3.0
        rec["position"][0, 0] += np.float32(0.25)
        rec["position"][0, 2] -= np.float32(0.125)
        rec["rotation"][0, 1] = np.float32(-35.0 + 90.0 + 10.0 * k)
        rec = gpu_lib.prepare(rec)                            # rotation changed: refresh the host-computed Euler matrices
        w.update_bodies(0, scenes.CLASS_PLAYER, 0, rec)
        o.poke(scenes.CLASS_PLAYER, 0, rec)
        run_and_compare(("poke", k), frames=20)
    assert total > 100, total
    w.close()


def test_undefined_pairs_never_fire_and_are_rejected_at_the_boundary(gpu_lib):
    """sphere-capsule and capsule-capsule are empty functions upstream.  Table-level queries are
    refused with PG_ERR_UNSUPPORTED; inside a step such a visit never fires: a sphere flies straight through a
    capsule player, neither is touched, no event."""
    from crux_b200.gpu import PgError
    sph = scenes.sphere_bodies(np.array([[0.0, 1.0, 0.6]], np.float32), np.array([[0.0, 0.0, -3.0]], np.float32))
    cap = gen.new_bodies(1)
    cap["collider_type"] = scenes.COLLIDER_CAPSULE if hasattr(scenes, "COLLIDER_CAPSULE") else 2
    cap["entity_type"] = 3
    cap["shape"][0, :3] = [0, 0.25, 0]; cap["shape"][0, 3:6] = [0, 1.75, 0]; cap["shape"][0, 6] = 0.25
    cap["has_node"] = 1; cap["has_parent"] = 1
    cap["node_xform"] = gen.node_xform(gpu_lib.node_transform, [0, 0, 0], [0, 0, 0], [1, 1, 1], np.eye(4, dtype=np.float32).reshape(16))
    for a, b in ((sph, cap), (cap, cap)):
        for call in (lambda: gpu_lib.pair_min_dist(a, b, 0.0), lambda: gpu_lib.pair_interval_collision(a, b, 0.0, DT),
                     lambda: gpu_lib.pair_narrow_phase(a, b, DT), lambda: gpu_lib.pair_resolve(a, b, 0.0, DT)):
            with pytest.raises(PgError, match="-4"):
                call()
        ev, a2, b2 = gpu_lib.pair_visit(a, b, DT)
        assert ev[0] == 0 and gen.same_bits(a2["position"], gpu_lib.prepare(a.copy())["position"]) and gen.same_bits(b2["velocity"], b["velocity"])
    # in a world: the sphere crosses the capsule's volume in ~20 frames; only gravity acts on it
    w = gpu_lib.GpuWorlds(1, 1, 1, 1, max_events_per_world=64)
    floor = scenes.plane_bodies(np.array([[0.0, 1.0, 0.0, -50.0]], np.float32))
    w.set_bodies(0, scenes.CLASS_STATIC, floor)
    w.set_bodies(0, scenes.CLASS_DYNAMIC, sph)
    w.set_bodies(0, scenes.CLASS_PLAYER, cap)
    w.step(DT, 30, refresh_nodes=True)
    assert len(w.events()) == 0
    d = w.get_bodies(0, scenes.CLASS_DYNAMIC)
    assert d["position"][0, 2] < -0.8 and d["velocity"][0, 0] == 0.0 and d["velocity"][0, 2] == -3.0
    w.close()


def test_statistics_come_from_the_step_epilogue_and_the_library_allreduce(gpu_lib, oracle_lib):
    """PgStats of a sphere batch are reduced in k_step_spheres' epilogue (no separate pass): compare them with
    the same aggregates taken from the downloaded state, after a step launch, after a chunked host step and for
    a state that was uploaded but not stepped (zero-step launch).  The library's own NCCL communicator (one
    rank here; bench.py runs it over 2-8 GPUs) must return the same vector."""
    from crux_b200 import shard
    nw, n, L = 37, 200, 9.0
    pos, vel = scenes.sphere_worlds(nw, n, L, first_world=900)
    st = scenes.plane_bodies(scenes.box_planes(L))
    w = gpu_lib.GpuWorlds(nw, len(st), n, 0, max_events_per_world=4096)
    w.set_spheres(0, st, pos, vel)

    def check(tag, contacts=None):
        s = w.stats()
        P = np.zeros((nw, n, 3), np.float32); V = np.zeros((nw, n, 3), np.float32); R = np.zeros((nw, n), np.uint8)
        w.download_state(0, nw, P, V, R)
        want = shard.stats_from_state(V, R)
        got = np.array([s[k] for k in gpu_lib.PG_STATS_FIELDS])
        assert got[0] == nw * n and got[7] == 0, tag
        assert np.allclose(got[[1, 2, 3, 4, 6]], want[[1, 2, 3, 4, 6]], rtol=1e-12, atol=1e-9), (tag, got, want)
        if contacts is not None:
            assert got[5] == contacts, (tag, got[5], contacts)
        return got

    check("fresh upload (zero-step launch)", contacts=0)
    w.step(DT, 40)
    got = check("after a step launch", contacts=len(w.events()))
    assert got[5] > 100
    launches = w.launches
    w.stats(); w.stats()
    assert w.launches == launches, "statistics of a stepped batch must not launch anything"
    s1 = w.stats_allreduce(None, want_host=True)
    assert np.array_equal(np.array(list(s1.values())), got)
    if gpu_lib.lib().pg_comm_nccl_version() > 0:
        comm = gpu_lib.Comm(gpu_lib.Comm.unique_id(), 0, 1, 0)
        s2 = w.stats_allreduce(comm, want_host=True)
        assert np.array_equal(np.array(list(s2.values())), got)
        assert w.last_step_ms() > 0.0
        comm.close()
    # chunked host pipeline: every chunk's launch adds its partials
    P = gpu_lib.PinnedArray((nw, n, 3), np.float32); V = gpu_lib.PinnedArray((nw, n, 3), np.float32); R = gpu_lib.PinnedArray((nw, n), np.uint8)
    w.download_state(0, nw, P.array, V.array, R.array)
    w.step_host(DT, 3, P.array, V.array, R.array)
    check("after step_host")
    # replace the state without stepping: the statistics follow
    V.array[:] = 0.0
    w.upload_state(0, nw, P.array, V.array, R.array)
    s = w.stats()
    assert s["kinetic"] == 0.0 and s["bodies"] == nw * n
    for a in (P, V, R):
        a.free()
    w.close()


def test_general_kernels_random_campaign(gpu_lib, oracle_lib):
    """Random mixed worlds (sizes, collider mixes, restitutions, time steps, with and without a player) through the
    speculate-and-commit kernel of game-sized worlds, in random chunk lengths: state, node matrices and event order
    bit-exact against the oracle (tests/random_campaign_mixed.py ran 300 cases of it)."""
    import random_campaign_mixed
    assert random_campaign_mixed.run(10, seed=77, steps=80, verbose=False) > 200
