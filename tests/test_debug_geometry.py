"""Debug-geometry interop: the numbers the reference's debug renderer
(src/physics/debug_renderer.c) uploads and draws, against what this library produces.

The fixture tests/golden/debug_geometry.npz was written by tests/make_golden.py from the UNMODIFIED
debug_renderer.c running behind a recording OpenGL stub (oracle/ref_debug_harness.c).  CPU tests pin the
fixture to a live run of that build (where oracle/_ref exists) and check the host-side mesh helpers; the
GPU tests compare the CUDA kernel's records bit for bit, through the C-ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import gen
import ora
from crux_b200 import scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = scenes.DT_60HZ


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12"):
        try:
            rt = C.CDLL(name)
        except OSError:
            continue
        rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        rt.cudaFree.argtypes = [C.c_void_p]
        return rt
    pytest.fail("the CUDA runtime library is not on this box")


def _fixture():
    return np.load(os.path.join(G, "debug_geometry.npz"))


def _same_records(a, b):
    assert len(a) == len(b)
    for name in ("kind", "cls", "index", "n_indices"):
        assert np.array_equal(a[name], b[name]), name
    bad = [k for k in range(len(a)) if not gen.same_bits(a["data"][k], b["data"][k])]
    assert not bad, (bad[:5], a["data"][bad[0]], b["data"][bad[0]])


# ------------------------------------------------------------------ CPU: fixture pinned to the reference, host helpers

def test_fixture_equals_live_reference_debug_renderer(oracle_lib):
    if not os.path.exists(ora.REF_DEBUG_SO):
        if not os.path.isdir(ora.REFERENCE_ROOT):
            pytest.skip("oracle/_ref/libcrux_ref_debug.so not built and /root/reference absent")
        ora.build_oracles()
    dbg = ora.RefDebugLib()
    z = _fixture()
    w = ora.World(dbg, z["gallery_S"], z["gallery_D"], z["gallery_P"])
    init_log = dbg.debug_init(w)
    recs, host = gen.debug_records_from_log(dbg.debug_render(w), z["gallery_S"], z["gallery_D"], z["gallery_P"])
    _same_records(recs, z["gallery_records"])
    assert sorted(host) == z["gallery_host_idx"].tolist()
    for k, m in zip(z["gallery_host_idx"], z["gallery_host_models"]):
        assert gen.same_bits(host[int(k)], m)
    assert len(init_log) == 2 * len(z["mesh_kind"])
    for k in range(len(z["mesh_kind"])):
        v, idx = init_log[2 * k][3], init_log[2 * k + 1][3]
        assert gen.same_bits(v, z["mesh_vertices"][k][:len(v)]) and np.array_equal(idx, z["mesh_indices"][k][:len(idx)])


def test_host_mesh_helpers_equal_the_reference_meshes():
    """pg_debug_sphere_mesh / capsule_mesh / box_mesh against what physics_debug_*_init uploaded (debug_renderer.c:311-639)."""
    from crux_b200 import gpu
    z = _fixture()
    seen = set()
    for k, kind in enumerate(z["mesh_kind"]):
        shape = z["mesh_shape"][k]
        if kind == scenes.COLLIDER_AABB:
            v, idx = gpu.debug_box_mesh(shape[:3], shape[3:6])
        elif kind == scenes.COLLIDER_SPHERE:
            v, idx = gpu.debug_sphere_mesh(shape[3])
        else:
            v, idx = gpu.debug_capsule_mesh(shape[:3], shape[3:6], shape[6])
        seen.add(int(kind))
        assert gen.same_bits(v.reshape(-1), z["mesh_vertices"][k][:v.size]), (k, kind)
        assert np.array_equal(idx, z["mesh_indices"][k][:idx.size]), (k, kind)
        assert not z["mesh_vertices"][k][v.size:].any() and not z["mesh_indices"][k][idx.size:].any()
    assert seen == {scenes.COLLIDER_AABB, scenes.COLLIDER_SPHERE, scenes.COLLIDER_CAPSULE}
    # tilted / flipped (rotate-by-pi branch) / degenerate capsules
    for k, shape in enumerate(z["capsule_shapes"]):
        v, idx = gpu.debug_capsule_mesh(shape[:3], shape[3:6], shape[6])
        assert gen.same_bits(v.reshape(-1), z["capsule_vertices"][k]), k
        assert np.array_equal(idx, z["capsule_indices"][k]), k


def test_host_capsule_model_equals_the_reference():
    """Capsules drawn without a scene node: model = T R_y R_x R_z S (debug_renderer.c:53-59, :109-116)."""
    from crux_b200 import gpu
    z = _fixture()
    order = np.concatenate([ora.from_core(ora.to_core(z["gallery_P"])), ora.from_core(ora.to_core(z["gallery_S"])),
                            ora.from_core(ora.to_core(z["gallery_D"]))])
    assert len(z["gallery_host_idx"]) >= 2
    for k, m in zip(z["gallery_host_idx"], z["gallery_host_models"]):
        assert gen.same_bits(gpu.debug_capsule_model(order[int(k):int(k) + 1]), m), k


# ------------------------------------------------------------------ GPU: the kernel, through the C-ABI

def _world(gpu_lib, S, D, P):
    w = gpu_lib.GpuWorlds(1, max(len(S), 1), max(len(D), 1), max(len(P), 1), max_events_per_world=4096)
    w.set_bodies(0, scenes.CLASS_STATIC, ora.from_core(ora.to_core(S)))
    w.set_bodies(0, scenes.CLASS_DYNAMIC, ora.from_core(ora.to_core(D)))
    w.set_bodies(0, scenes.CLASS_PLAYER, ora.from_core(ora.to_core(P)))
    return w


@pytest.mark.gpu
def test_debug_geometry_gallery_bit_exact(gpu_lib):
    """Every drawable kind (AABB corners through the node, sphere models, node-driven capsules) equals what the
    reference handed to glBufferSubData / shader_set_mat4; into host memory and into a caller-owned DEVICE buffer
    (the stand-in for a mapped GL buffer)."""
    z = _fixture()
    w = _world(gpu_lib, z["gallery_S"], z["gallery_D"], z["gallery_P"])
    got = w.debug_geometry(0)
    _same_records(got, z["gallery_records"])
    assert set(got["kind"].tolist()) == {0, 1, 2, 3, 4}
    # a device buffer the CALLER owns, from the CUDA runtime directly (no torch: the buffer is what a GL host would map)
    rt = _cudart()
    nbytes = len(got) * scenes.DEBUG_DTYPE.itemsize
    dev = C.c_void_p()
    assert rt.cudaMalloc(C.byref(dev), C.c_size_t(nbytes)) == 0
    assert rt.cudaMemset(dev, 0, C.c_size_t(nbytes)) == 0
    launches = w.launches
    n = w.debug_geometry(0, device_ptr=dev.value, max_records=len(got))
    assert n == len(got) and w.launches == launches + 1
    w.sync()
    back = np.zeros(len(got), scenes.DEBUG_DTYPE)
    assert rt.cudaMemcpy(back.ctypes.data_as(C.c_void_p), dev, C.c_size_t(nbytes), 2) == 0      # cudaMemcpyDeviceToHost
    _same_records(back, z["gallery_records"])
    with pytest.raises(gpu_lib.PgError):
        w.debug_geometry(0, device_ptr=dev.value, max_records=len(got) - 1)
    assert rt.cudaFree(dev) == 0
    w.close()


@pytest.mark.gpu
def test_debug_geometry_follows_the_stepped_world(gpu_lib):
    """The mixed world (planes, static and dynamic AABBs, a capsule player) after 40 frames with node refresh:
    the kernel reads the RESIDENT state, no read-back of bodies in between."""
    z = _fixture()
    m = np.load(os.path.join(G, "mixed_world.npz"))
    w = _world(gpu_lib, m["statics"], m["dynamics"], m["players"])
    w.step(DT, int(z["mixed_steps"]), refresh_nodes=True)
    _same_records(w.debug_geometry(0), z["mixed_records"])
    assert (z["mixed_records"]["kind"] == scenes.DEBUG_AABB_LINES).sum() > 20
    w.close()


@pytest.mark.gpu
def test_debug_geometry_sphere_batch(gpu_lib):
    """Sphere batches keep float4 state, not records: a dynamic sphere's model is translate(position + 0) (:184-193),
    planes draw nothing."""
    nw, n = 3, 64
    pos, vel = scenes.sphere_worlds(nw, n, 10.0)
    st = scenes.plane_bodies(scenes.box_planes(10.0))
    w = gpu_lib.GpuWorlds(nw, len(st), n, 0)
    w.set_spheres(0, st, pos, vel, radius=0.15, restitution=1.0)
    w.step(DT, 7)
    p = np.zeros((nw, n, 3), np.float32); v = np.zeros_like(p)
    w.download_state(0, nw, p, v)
    for world in (0, 2):
        got = w.debug_geometry(world)
        assert len(got) == len(st) + n
        assert (got["kind"][:len(st)] == scenes.DEBUG_NONE).all() and (got["cls"][:len(st)] == scenes.CLASS_STATIC).all()
        d = got[len(st):]
        assert (d["kind"] == scenes.DEBUG_SPHERE_MODEL).all() and (d["n_indices"] == 2280).all()
        want = np.zeros((n, 24), np.float32)
        want[:, 0] = want[:, 5] = want[:, 10] = want[:, 15] = 1.0
        want[:, 12:15] = p[world] + np.float32(0.0)
        assert gen.same_bits(d["data"], want)
    w.close()
