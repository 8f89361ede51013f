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
