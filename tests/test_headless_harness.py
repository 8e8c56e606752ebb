"""The C harness (tools/rt_headless.c: the reference's command line on top of the
C ABI, host code in C) against the oracle: reference flags --scene/--threads/
--init-scale, progressive passes 16->1 then continued accumulation, column
layout of --threads, key replay, and the screenshot quantise+flip rule."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bindings

HARNESS = os.path.join(ROOT, "tools", "rt_headless")
SCENE = os.path.join(bindings.ASSETS, "scene_0.txt")
SKYDIR = os.path.join(bindings.ASSETS, "skybox")

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.skipif(not (os.path.exists(HARNESS) and os.path.exists(SCENE) and os.path.exists(os.path.join(SKYDIR, "front.jpg"))),
                    reason="harness or staged reference assets missing")
def test_harness_progressive_frames_match_oracle(tmp_path, port, real_sky):
    from ray_tracing_b200 import host

    W, H, T = 256, 144, 4
    raw, ppm = tmp_path / "f.raw", tmp_path / "f.ppm"
    out = subprocess.run([HARNESS, "--scene", SCENE, "--threads", str(T), "--init-scale", "16", "--width", str(W), "--height", str(H),
                          "--frames", "7", "--keys", "W", "--skybox", SKYDIR, "--dump-f32", str(raw), "--dump", str(ppm)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    info = json.loads(out.stdout.strip().splitlines()[-1])
    assert info["frames"] == 7 and info["rays"] > 0
    got = np.fromfile(raw, np.float32).reshape(H, W, 3)

    # oracle: key W moves the camera before frame 0 (main.c:536-540), then passes at 16,8,4,2,1,1,1
    host.camera_reset()
    host.move_camera(host.RT_UP, 0.5)
    cam = host.camera_snapshot().as_dict()
    host.camera_reset()
    objs = host.parse_scene_file(SCENE)
    world = port.world(objs, real_sky, cam)
    acc = np.zeros((H, W, 3), np.float32)
    count = np.float32(0)
    for p, s in enumerate((16, 8, 4, 2, 1, 1, 1)):
        data, _ = port.render(world, W, H, s, T, p)
        port.accumulate(acc, data, s)
        count = np.float32(count + np.float32(1.0) / np.float32(s * s))
    want = port.resolve(acc, count)
    assert np.array_equal(bits(got), bits(want))
    assert abs(info["accum_count"] - float(count)) < 1e-3      # printed with 4 decimals

    # screenshot rule: (uint8_t)(x*255), flipped vertically, P6 (PNG covered in test_camera_host.py)
    with open(ppm, "rb") as f:
        assert f.readline() == b"P6\n" and f.readline() == f"{W} {H}\n".encode() and f.readline() == b"255\n"
        px = np.frombuffer(f.read(), np.uint8).reshape(H, W, 3)
    assert np.array_equal(px, host.quantize_frame(want)[::-1])


def test_harness_rejects_bad_arguments():
    if not os.path.exists(HARNESS):
        pytest.skip("harness not built")
    r = subprocess.run([HARNESS, "--threads", "4"], capture_output=True, text=True)
    assert r.returncode != 0 and "No scene specified" in r.stderr
    r = subprocess.run([HARNESS, "--scene", "x.txt"], capture_output=True, text=True)
    assert r.returncode != 0 and "Missing --threads" in r.stderr
    r = subprocess.run([HARNESS, "--scene", "x.txt", "--threads", "4", "--init-scale", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid value for --init-scale" in r.stderr
