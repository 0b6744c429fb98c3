"""Drop-in test of the plugin boundary: the headless host `nans` loads the new `nans.so`, which
steps the reference's Init scene on the GPU through SimUpdateAndRender — scripted aim + shot
included, FREE-RUNNING for 1000 frames — and the trajectory must equal, bit for bit, the one the
reference's own nans.so produced through the same entry point (tests/golden/demo_traj.npz).
A second run forces a dlclose/dlopen hot reload mid-simulation."""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_bit_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nans_projekat_b200", "host", "nans")


def read_dump(path, nb):
    rec = 4 * 4 + nb * 12 * 4 + 8 * 4 + 2 * 64
    raw = np.fromfile(path, np.uint8)
    assert raw.size % rec == 0 and raw.size > 0
    raw = raw.reshape(-1, rec)
    hdr = raw[:, :16].copy().view(np.int32)
    body = raw[:, 16:16 + nb * 48].copy().view(np.float32).reshape(-1, 4, nb, 3)
    o = 16 + nb * 48
    cam = raw[:, o:o + 32].copy().view(np.float32)
    view = raw[:, o + 32:o + 96].copy().view(np.float32)
    proj = raw[:, o + 96:o + 160].copy().view(np.float32)
    return hdr, body, cam, view, proj


@pytest.mark.parametrize("reload_at", [-1, 500])
def test_demo_scene_through_plugin(tmp_path, golden_dir, reload_at):
    if not os.path.exists(HOST):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    z = np.load(os.path.join(golden_dir, "demo_traj.npz"))
    frames = len(z["pos"])
    dump = tmp_path / "traj.bin"
    cmd = [HOST, "--frames", str(frames), "--dt", "0.016666668", "--script", "demo", "--dump", str(dump), "--quiet"]
    if reload_at >= 0:
        cmd += ["--reload-at", str(reload_at)]
    env = dict(os.environ)
    env.pop("NANS_SCENE", None)
    subprocess.check_call(cmd, env=env)
    hdr, body, cam, view, proj = read_dump(dump, 5)
    assert len(hdr) == frames
    assert np.array_equal(hdr[:, 2], z["ncontacts"]), "contact count per frame"
    for i, f in enumerate(("pos", "ang", "vel", "angvel")):
        assert_bit_equal(body[:, i], z[f], f"free-running {f} over {frames} frames")
    assert_bit_equal(cam, z["camera"], "camera position/front/yaw/pitch")
    # the shot happened and moved the sphere
    assert np.abs(z["vel"][202, 4]).max() > 10 and np.abs(body[202, 2, 4]).max() > 10
    # View / Projection written back to the host every frame (finite, perspective shape)
    assert np.isfinite(view).all() and np.isfinite(proj).all()
    assert np.allclose(proj[:, 11], -1.0) and np.allclose(proj[:, 15], 0.0)
