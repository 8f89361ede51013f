"""Config C1 through the actual drop-in boundary: the reference-named C API (physics_world_create,
physics_add_body, physics_add_player, physics_step, ...) implemented by libcrux_physics.so on the
GPU, driven by the headless scene loader on the reference's own scene data, against dumps produced
by the SAME driver linked with the unmodified reference physics (oracle/_ref/crux_headless_ref,
tests/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "crux_b200", "lib", "crux_headless")
SC = os.path.join(ROOT, "tests", "golden", "scenes")


def parse_dump(blob):
    hdr = np.frombuffer(blob[:16], np.int32)
    nb = int(hdr[0] + hdr[1] + hdr[2])
    bodies = np.frombuffer(blob[16:16 + nb * 92], np.float32).reshape(nb, 23)
    events = np.frombuffer(blob[16 + nb * 92:], np.int32).reshape(-1, 3)
    return hdr, bodies, events


@pytest.mark.parametrize("name,player,steps", [("bouncehouse", True, 1000), ("bouncehouse3", False, 600), ("scene_sphere", False, 600)])
@pytest.mark.parametrize("device_steps", [False, True])
def test_scene_file_through_reference_api(tmp_path, gpu_lib, name, player, steps, device_steps):
    if not os.path.exists(EXE):
        pytest.fail("crux_headless not built (python -m crux_b200.build)")
    out = tmp_path / "gpu.bin"
    cmd = [EXE, os.path.join(SC, name + ".json"), str(steps), str(out)]
    if player:
        cmd.append("--player")
    if device_steps:
        cmd.append("--device-steps")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # dirty-range upload: only records the HOST changed since the device last saw
    # them travel.  Frame by frame the scene rewrites the node matrix of every body that moved
    # (scene_node_update), so those go up again and the static level never does; with the frames run on
    # the device there is exactly one upload.
    up = [int(l.split(":")[1]) for l in r.stdout.splitlines() if l.startswith("uploaded body records")]
    assert up, r.stdout
    ns, nd, npl = (int(x) for x in np.frombuffer(open(out, "rb").read()[:12], np.int32))
    if device_steps:
        assert up[0] == ns + nd + npl, (up, ns, nd, npl)
    else:
        assert ns + nd + npl <= up[0] <= ns + steps * (nd + npl), (up, ns, nd, npl)
    got = open(out, "rb").read()
    want = open(os.path.join(SC, f"{name}_{steps}.bin"), "rb").read()
    if got != want:
        hg, bg, eg = parse_dump(got)
        hw, bw, ew = parse_dump(want)
        assert list(hg) == list(hw), (list(hg), list(hw))
        bad = np.argwhere(bg.view(np.uint32) != bw.view(np.uint32))
        assert len(bad) == 0, f"first differing (body, field): {bad[:5].tolist()} got {bg[bad[0][0]]} want {bw[bad[0][0]]}"
        assert np.array_equal(eg, ew)
    assert got == want


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def _assert_same_dump(got, want):
    if got != want:
        hg, bg, eg = parse_dump(got)
        hw, bw, ew = parse_dump(want)
        assert list(hg) == list(hw), (list(hg), list(hw))
        bad = np.argwhere(bg.view(np.uint32) != bw.view(np.uint32))
        assert len(bad) == 0, f"first differing (body, field): {bad[:5].tolist()} got {bg[bad[0][0]]} want {bw[bad[0][0]]}"
        assert np.array_equal(eg, ew)
    assert got == want


def test_item_pickup_removes_bodies_mid_game(tmp_path, gpu_lib):
    """three dynamic bodies are ITEM entities; the player (walked under them by position
    writes) triggers EVENT_PLAYER_ITEM_PICKUP, whose handling removes the item's body with physics_remove_body
    (event.c:118-125 -> scene.c:1203 -> world.c:147-172, swap-and-pop: the solve order of the frames that
    follow changes); two more removals hit the first and the last dynamic slot.  400 frames, byte-identical to
    the same driver linked with the unmodified reference physics."""
    out = tmp_path / "gpu.bin"
    text = _run([EXE, os.path.join(SC, "bouncehouse.json"), "400", str(out), "--player", "--script", "pickup"])
    assert "script pickup: 5 bodies removed" in text, text
    got = open(out, "rb").read()
    hdr, bodies, events = parse_dump(got)
    assert list(hdr[:3]) == [6, 3, 1] and int((events[:, 0] == 1).sum()) == 3      # three pickups happened
    _assert_same_dump(got, open(os.path.join(SC, "bouncehouse_pickup_400.bin"), "rb").read())


@pytest.mark.parametrize("scene", ["bouncehouse", "plates_only"])
def test_player_pokes_between_frames(tmp_path, gpu_lib, scene):
    """the player write path of player.c -- jump (at_rest = false, velocity[1] = 3;
    player.c:160-173), walking (position writes + scene_node_update; player.c:11-50), turning (rotation write;
    player.c:155-157) -- between physics_step calls, 400 frames, byte-identical to the reference-linked twin.
    In the scene where nothing but the player moves the dirty-range upload must send exactly the player's
    record in a frame that poked (or moved) it, and nothing at all once it rests unpoked."""
    out = tmp_path / "gpu.bin"
    text = _run([EXE, os.path.join(SC, scene + ".json"), "400", str(out), "--player", "--script", "jump"])
    assert "4 jumps" in text, text
    _assert_same_dump(open(out, "rb").read(), open(os.path.join(SC, f"{scene}_jump_400.bin"), "rb").read())
    up = {int(l.split()[1]): int(l.split()[3]) for l in text.splitlines() if l.startswith("frame ") and " uploaded " in l}
    assert len(up) == 400
    if scene == "plates_only":
        assert up[0] == 7                                   # first frame: the six plates and the player
        assert all(v <= 1 for k, v in up.items() if k > 0), [kv for kv in up.items() if kv[1] > 1][:5]
        for k in (30, 120, 121, 300, 60, 75, 200):          # frames whose script poked the player record
            assert up[k] == 1, (k, up[k])
        assert sum(1 for k, v in up.items() if k > 0 and v == 0) > 100      # resting and unpoked: nothing travels


@pytest.mark.gpu
def test_debug_geometry_through_the_reference_named_api(tmp_path, gpu_lib):
    """physics_debug_geometry (the drop-in's counterpart of the reads physics_debug_render makes, debug_renderer.c:11-213)
    on scenes/bouncehouse.json + player after 200 frames and one more host-side poke of the player: the poke must reach
    the device before the kernel runs (dirty-range upload of a zero-step frame), the player's capsule record carries the
    scene node's world transform bit for bit, every AABB record is a box around its node."""
    from crux_b200 import scenes
    out, dbg = tmp_path / "gpu.bin", tmp_path / "debug.bin"
    r = _run([EXE, os.path.join(SC, "bouncehouse.json"), "200", str(out), "--player", "--debug-geometry", str(dbg)])
    hdr, bodies, _ = parse_dump(open(out, "rb").read())
    ns, nd, npl = (int(x) for x in hdr[:3])
    recs = np.frombuffer(open(dbg, "rb").read(), dtype=scenes.DEBUG_DTYPE)
    assert len(recs) == ns + nd + npl and f"debug geometry: {len(recs)} records" in r
    # draw order: players, statics, dynamics; the dump is statics, dynamics, players
    assert recs["cls"].tolist() == [2] * npl + [0] * ns + [1] * nd
    player = bodies[ns + nd]
    assert recs["kind"][0] == scenes.DEBUG_CAPSULE_MODEL and recs["n_indices"][0] == 4920
    assert np.array_equal(recs["data"][0][:16].view(np.uint32), player[7:23].view(np.uint32))
    assert recs["data"][0][12] == player[0]                      # ... which includes the poke (position written into the node)
    dumped = np.concatenate([bodies[:ns], bodies[ns:ns + nd]])
    boxes = recs[npl:]
    assert (boxes["kind"] == scenes.DEBUG_AABB_LINES).all() and (boxes["n_indices"] == 24).all()
    corners = boxes["data"].reshape(-1, 8, 3)
    assert (corners[:, 0] >= corners[:, 4]).all()                # corner 0 = centre + extents, corner 4 = centre - extents
    centre = 0.5 * (corners[:, 0] + corners[:, 4])
    assert np.abs(centre - dumped[:, 19:22]).max() < 2.0         # the box sits around its node's translation
