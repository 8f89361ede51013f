"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled in this container
(oracle/_ref/libcrux_ref.so, built by oracle/Makefile from /root/reference/src/physics/*.c).

    python tests/make_golden.py

The fixtures travel with the repo so that boxes without /root/reference (the GPU box) can still pin
the plain-C restatement and the CUDA path to the reference's own outputs:
  pairs_kat.npz        per-pair known answers for the 7 live collider pairs: distance, interval
                       gate, narrowphase, resolver and full-visit outputs on seeded random and
                       adversarial inputs (touching, NaN root, zero relative velocity, behind-plane)
  spheres_256.npz      256-sphere bouncehouse world (world 0): state after
                       1, 10, 100, 1000 steps + the contact-event sequence of the first 100 steps
  mixed_world.npz      planes + static AABBs + dynamic AABBs (incl. an item) + capsule player with
                       scene nodes refreshed every step: states after 25/50/150 steps + events
  aabb_unity.npz       inputs/expected outputs of the reference's own Unity test
                       (test/physics/test_aabb.c:24-141)
  debug_geometry.npz   what the UNMODIFIED debug renderer (src/physics/debug_renderer.c, behind the recording GL stub
                       of oracle/ref_debug_harness.c) uploads and draws: per-body corner vertices / model matrices
                       of a gallery of every drawable kind and of the mixed world after 40 stepped frames, and the
                       init-time sphere / capsule / box meshes
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import gen  # noqa: E402
import ora  # noqa: E402
from crux_b200 import scenes  # noqa: E402

OUT = os.path.join(HERE, "golden")


def oracle_node_transform(orc):
    import ctypes as C

    def f(pos, rot, scale, parent=None):
        out = np.zeros(16, np.float32)
        p = np.ascontiguousarray(pos, np.float32); r = np.ascontiguousarray(rot, np.float32)
        s = np.ascontiguousarray(scale, np.float32)
        par = None if parent is None else np.ascontiguousarray(parent, np.float32)
        orc.dll.ora_node_transform(p.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                   None if par is None else par.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out
    return f


def adversarial_pairs():
    """Hand-made sphere/plane records for the narrowphase quirks listed in SURVEY.md a23/a24."""
    A, B = [], []

    def sph(p, v, r=0.15):
        return scenes.sphere_bodies(np.array([p], np.float32), np.array([v], np.float32), radius=r)[0]

    def pln(n, d):
        return scenes.plane_bodies(np.array([[*n, d]], np.float32))[0]
    A.append(sph([0, 1, 0], [0, 0, 0])); B.append(sph([0.3, 1, 0], [0, 0, 0]))            # touching, a = 0 -> NaN root
    A.append(sph([0, 1, 0], [1, 0, 0])); B.append(sph([0.30001, 1, 0], [-1, 0, 0]))        # head-on, just apart
    A.append(sph([0, 1, 0], [1, 0, 0])); B.append(sph([0.29, 1, 0], [1, 0, 0]))            # overlapping, same velocity
    A.append(sph([0, 1, 0], [-1, 0, 0])); B.append(sph([0.4, 1, 0], [1, 0, 0]))            # receding -> negative root
    A.append(sph([0, 1, 0], [3, 0, 0])); B.append(sph([0.35, 1.29, 0], [0, 0, 0]))         # grazing, d < 0
    A.append(sph([0, 0.1, 0], [0, -1, 0])); B.append(pln([0, 1, 0], 0))                    # penetrating floor
    A.append(sph([0, -0.5, 0], [0, 1, 0])); B.append(pln([0, 1, 0], 0))                    # behind the plane
    A.append(sph([0, 0.15, 0], [1, 0, 0])); B.append(pln([0, 1, 0], 0))                    # resting contact, tangential motion
    A.append(sph([0, 0.150024, 0], [0.75, 0, 11.65])); B.append(pln([0, 1, 0], 0))         # the 2^depth case
    A.append(sph([0, 0.2, 0], [0, -0.3, 0])); B.append(pln([0, 1, 0], 0))                  # slow approach -> resting rule
    return np.array(A, dtype=scenes.BODY_DTYPE), np.array(B, dtype=scenes.BODY_DTYPE)


def main():
    ora.build_oracles()
    ref = ora.RefLib()
    orc = ora.OracleLib()
    nt = oracle_node_transform(orc)
    os.makedirs(OUT, exist_ok=True)
    dt = gen.DT

    # ---- per-pair KATs
    rng = np.random.default_rng(20261019)
    A, B = gen.rand_pairs(rng, 700, nt)
    A2, B2 = adversarial_pairs()
    A = np.concatenate([A, A2]); B = np.concatenate([B, B2])
    n = len(A)
    t = rng.uniform(0, dt, n).astype(np.float32)
    out = dict(A=A, B=B, t=t, dist=np.zeros(n, np.float32), maxmove=np.zeros(n, np.float32),
               interval=np.zeros(n, np.int32), colliding=np.zeros(n, np.int32), hit_time=np.zeros(n, np.float32),
               visit=np.zeros(n, np.int32), A_after=scenes_core(n), B_after=scenes_core(n),
               A_resolved=scenes_core(n), B_resolved=scenes_core(n), resolve_time=np.zeros(n, np.float32))
    for k in range(n):
        a, b = A[k:k + 1], B[k:k + 1]
        out["dist"][k] = ref.k_min_dist(a, b, t[k])
        out["maxmove"][k] = ref.k_max_move(a, 0.0, dt)
        out["interval"][k] = ref.k_interval(a, b, 0.0, dt)[0]
        c, ht = ref.k_narrow(a, b, dt)
        out["colliding"][k], out["hit_time"][k] = c, ht
        ev, aa, bb = ref.k_visit(a, b, dt)
        out["visit"][k] = ev; out["A_after"][k] = aa[0]; out["B_after"][k] = bb[0]
        rt = ht if ht >= 0 else np.float32(0.004)
        out["resolve_time"][k] = rt
        _, aa, bb = ref.k_resolve(a, b, rt, dt)
        out["A_resolved"][k] = aa[0]; out["B_resolved"][k] = bb[0]
    np.savez_compressed(os.path.join(OUT, "pairs_kat.npz"), **out)
    print("pairs_kat:", n, "pairs,", int(out["visit"].sum()), "visits fire")

    # ---- 256-sphere world trajectory
    st, dy = scenes.bouncehouse_world(0)
    w = ora.World(ref, st, dy)
    snaps = {}
    done = 0
    events100 = None
    for target in (1, 10, 100, 1000):
        w.step(scenes.DT_60HZ, target - done)
        done = target
        snaps[f"state_{target}"] = w.read(1)
        if target == 100:
            events100 = w.events(clear=False)
    ev_all = w.events()
    np.savez_compressed(os.path.join(OUT, "spheres_256.npz"), events_first100=events100, n_events_1000=len(ev_all), **snaps)
    print("spheres_256: events in 1000 steps =", len(ev_all))

    # ---- mixed world with scene nodes
    rng = np.random.default_rng(5)
    s_, d_, p_ = gen.mixed_world(rng, nt, with_player=True)
    w = ora.World(ref, s_, d_, p_)
    mixed = dict(statics=s_, dynamics=d_, players=p_)
    done = 0
    for target in (25, 50, 150):
        w.step(scenes.DT_60HZ, target - done, 1)
        done = target
        for cls, name in ((0, "S"), (1, "D"), (2, "P")):
            mixed[f"{name}_{target}"] = w.read(cls)
    mixed["events"] = w.events()
    np.savez_compressed(os.path.join(OUT, "mixed_world.npz"), **mixed)
    print("mixed_world: events =", len(mixed["events"]))

    # ---- the reference's own Unity vectors (test/physics/test_aabb.c)
    unity = dict(
        aabb_true_a=np.array([1.5, 1.5, 1.5, 0.5, 0.5, 0.5], np.float32), aabb_true_b=np.array([2.0, 2.0, 2.0, 0.5, 0.5, 0.5], np.float32),
        aabb_false_a=np.array([1.5, 1.5, 1.5, 0.5, 0.5, 0.5], np.float32), aabb_false_b=np.array([3.5, 3.5, 3.5, 0.5, 0.5, 0.5], np.float32),
        plane_box=np.array([1.5, 1.5, 1.5, 0.5, 0.5, 0.5], np.float32),
        plane_true_1=np.array([0.0, 1.0, 0.0, 1.5], np.float32), plane_true_2=np.array([0.577, 0.577, 0.577, 2.0], np.float32),
        plane_false_1=np.array([0.0, 1.0, 0.0, 0.0], np.float32), plane_false_2=np.array([0.577, 0.577, 0.577, 10.0], np.float32),
        update_src=np.array([0, 0, 0, 1, 1, 1], np.float32), update_translation=np.array([5, 0, 0], np.float32),
        update_scale=np.array([1, 1, 1], np.float32), update_rot_y_deg=np.float32(90.0),
        update_expected=np.array([5, 0, 0, 1, 1, 1], np.float32), update_tol=np.float32(1e-4))
    np.savez_compressed(os.path.join(OUT, "aabb_unity.npz"), **unity)
    make_debug_fixture(nt)
    print("wrote", sorted(os.listdir(OUT)))


def make_debug_fixture(nt):
    """tests/golden/debug_geometry.npz from the reference's own debug renderer (see the module docstring)."""
    dbg = ora.RefDebugLib()
    out = {}
    # (1) gallery: every drawable kind, not stepped
    rng = np.random.default_rng(77)
    S, D, P = gen.debug_gallery(rng, nt)
    w = ora.World(dbg, S, D, P)
    init_log = dbg.debug_init(w)
    recs, host_models = gen.debug_records_from_log(dbg.debug_render(w), S, D, P)
    out.update(gallery_S=S, gallery_D=D, gallery_P=P, gallery_records=recs,
               gallery_host_idx=np.array(sorted(host_models), np.int32),
               gallery_host_models=np.array([host_models[k] for k in sorted(host_models)], np.float32).reshape(-1, 16))
    # init-time meshes in draw order: (vertex data, index data) per drawable body
    order = [P, S, D]
    bodies = [b for arr in order for b in arr if int(b["collider_type"]) != scenes.COLLIDER_PLANE]
    assert len(init_log) == 2 * len(bodies), (len(init_log), len(bodies))
    mesh_kind, mesh_shape, mesh_v, mesh_i = [], [], [], []
    for k, b in enumerate(bodies):
        (wv, cv, _, v), (wi, ci, _, idx) = init_log[2 * k], init_log[2 * k + 1]
        assert wv == dbg.VERTEX_DATA and wi == dbg.INDEX_DATA
        mesh_kind.append(int(b["collider_type"])); mesh_shape.append(np.array(b["shape"], np.float32))
        vv = np.zeros(2646, np.float32); vv[:len(v)] = v
        ii = np.zeros(4920, np.uint32); ii[:len(idx)] = idx
        mesh_v.append(vv); mesh_i.append(ii)
    out.update(mesh_kind=np.array(mesh_kind, np.int32), mesh_shape=np.array(mesh_shape, np.float32),
               mesh_vertices=np.array(mesh_v), mesh_indices=np.array(mesh_i))
    # a few more capsule meshes: tilted, flipped (axis = -z: the rotate-by-pi branch), degenerate (A == B)
    rng = np.random.default_rng(78)
    segs = [((0, 0, 0), (0, 0, 1)), ((0, 0, 1), (0, 0, -1)), ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((1, 0, 0), (-1, 2, 0.5))]
    segs += [(tuple(rng.uniform(-1, 1, 3)), tuple(rng.uniform(-1, 1, 3))) for _ in range(4)]
    caps = gen.new_bodies(len(segs))
    caps["collider_type"] = scenes.COLLIDER_CAPSULE
    for k, (a, b) in enumerate(segs):
        caps["shape"][k, :3] = a; caps["shape"][k, 3:6] = b; caps["shape"][k, 6] = np.float32(0.2 + 0.1 * k)
    w2 = ora.World(dbg, caps, None, None)
    log2 = dbg.debug_init(w2)
    out.update(capsule_shapes=np.array(caps["shape"], np.float32),
               capsule_vertices=np.array([log2[2 * k][3] for k in range(len(segs))], np.float32),
               capsule_indices=np.array([log2[2 * k + 1][3] for k in range(len(segs))], np.uint32))
    # (2) the mixed world of mixed_world.npz (seed 5), drawn after 40 stepped frames with node refresh
    rng = np.random.default_rng(5)
    s_, d_, p_ = gen.mixed_world(rng, nt, with_player=True)
    w3 = ora.World(dbg, s_, d_, p_)
    w3.step(scenes.DT_60HZ, 40, 1)
    recs3, hm3 = gen.debug_records_from_log(dbg.debug_render(w3), w3.read(0), w3.read(1), w3.read(2))
    assert not hm3
    out.update(mixed_steps=np.int32(40), mixed_records=recs3)
    np.savez_compressed(os.path.join(OUT, "debug_geometry.npz"), **out)
    kinds = np.bincount(recs["kind"], minlength=5)
    print("debug_geometry: gallery kinds", kinds.tolist(), "meshes", len(bodies), "mixed records", len(recs3))


def scenes_core(n):
    return np.zeros(n, dtype=scenes.CORE_DTYPE)


if __name__ == "__main__":
    main()


# ---- scene-file fixtures (config C1): reduced copies of the reference's scene DATA + dumps of the
# reference-linked headless driver (oracle/_ref/crux_headless_ref)
SCENES = [("bouncehouse", True, 1000), ("bouncehouse3", False, 600), ("scene_sphere", False, 600)]
# the game layer's write paths (crux_headless --script): item pickup -> physics_remove_body, player pokes
SCRIPTED = [("bouncehouse", "pickup", 400), ("bouncehouse", "jump", 400), ("plates_only", "jump", 400)]


def _reduce(node):
    keep = ("position", "rotation", "scale", "velocity", "collider", "entity_type", "components")
    out = {k: node[k] for k in keep if k in node}
    if "components" in out:
        out["components"] = []                      # marks the Gen-3 schema; contents are not physics
    if "children" in node:
        out["children"] = [_reduce(c) for c in node["children"]]
    return out


def make_scene_fixtures():
    import json
    import subprocess
    out_dir = os.path.join(OUT, "scenes")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "crux_headless_ref")
    for name, player, steps in SCENES:
        src = os.path.join(ora.REFERENCE_ROOT, "scenes", name + ".json")
        doc = json.load(open(src))
        red = {}
        if "nodes" in doc:
            red["nodes"] = _reduce(doc["nodes"])
        if "meshes" in doc:
            red["meshes"] = {k: [_reduce(m) for m in v] for k, v in doc["meshes"].items()}
        dst = os.path.join(out_dir, name + ".json")
        json.dump(red, open(dst, "w"), indent=1)
        args = ["--player"] if player else []
        full = os.path.join(out_dir, f"{name}_{steps}_fromfull.bin")
        subprocess.run([exe, src, str(steps), full] + args, check=True)
        subprocess.run([exe, dst, str(steps), os.path.join(out_dir, f"{name}_{steps}.bin")] + args, check=True)
        assert open(full, "rb").read() == open(os.path.join(out_dir, f"{name}_{steps}.bin"), "rb").read(), "reduction changed the workload"
        os.remove(full)
    # our own reduced scene: the six plates of bouncehouse.json and nothing that moves but the player
    doc = json.load(open(os.path.join(out_dir, "bouncehouse.json")))

    def _static_only(node):
        kids = [_static_only(c) for c in node.get("children", [])
                if not (c.get("collider") and c["collider"].get("dynamic"))]
        out = {k: v for k, v in node.items() if k != "children"}
        if kids:
            out["children"] = kids
        return out
    json.dump({"nodes": _static_only(doc["nodes"])}, open(os.path.join(out_dir, "plates_only.json"), "w"), indent=1)
    for name, script, steps in SCRIPTED:
        subprocess.run([exe, os.path.join(out_dir, name + ".json"), str(steps), os.path.join(out_dir, f"{name}_{script}_{steps}.bin"),
                        "--player", "--script", script], check=True)


if __name__ == "__main__":
    make_scene_fixtures()
