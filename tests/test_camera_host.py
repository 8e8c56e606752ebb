"""Camera pose mutators (ray_tracing_b200/csrc/camera_host.c) against the
reference's camera.c:37-93 and the quantisation rule of main.c:666-670."""
import numpy as np

from ray_tracing_b200 import host


def snap():
    c = host.camera_snapshot()
    return np.array([*c.pos, *c.front, *c.up, c.fov], np.float32)


def ref_snap(ref):
    c = ref.get_camera()
    return np.concatenate([c["pos"], c["front"], c["up"], [c["fov"]]]).astype(np.float32)


def test_default_pose():
    host.camera_reset()
    assert snap().tolist() == [5, 5, 5, -1, -1, -1, 0, 1, 0, 30]
    assert host.get_camera_pos() == (5.0, 5.0, 5.0)


def test_first_mouse_event_snaps_front():
    host.camera_reset()
    host.rotate_camera(123.0, 456.0)
    c = host.camera_snapshot()
    # yaw -90, pitch 0 (camera.c:42-78): front ~ (-4e-8, 0, -1)
    assert abs(c.front[0]) < 1e-6 and c.front[1] == 0.0 and c.front[2] == -1.0


def test_mutator_sequences_match_reference(ref_pixel):
    rng = np.random.default_rng(9)
    host.camera_reset()
    ref_pixel.reset_camera()
    for step in range(300):
        kind = int(rng.integers(0, 3))
        if kind == 0:
            mx, my = float(rng.uniform(0, 1600)), float(rng.uniform(-4000, 4000))
            host.rotate_camera(mx, my)
            ref_pixel.rotate_camera(mx, my)
        else:
            d, sp = int(rng.integers(0, 4)), float(rng.choice([0.5, 0.25, 1.0]))
            host.move_camera(d, sp)
            ref_pixel.move_camera(d, sp)
        assert np.array_equal(snap().view(np.uint32), ref_snap(ref_pixel).view(np.uint32)), step
    host.camera_reset()
    ref_pixel.reset_camera()


def test_quantize_rule():
    x = np.array([[0.0, 0.5, 1.0], [0.999999, 0.003921, 0.00392157]], np.float32)
    q = host.quantize_frame(x)
    want = (x * np.float32(255)).astype(np.uint8)   # truncation
    assert np.array_equal(q, want)
    assert q[0].tolist() == [0, 127, 255]


def test_save_screenshot_png_and_ppm(tmp_path):
    """screenshot() pixel path (main.c:637-681): quantise, flip vertically, PNG."""
    rng = np.random.default_rng(3)
    frame = rng.uniform(0, 1, (37, 53, 3)).astype(np.float32)
    frame[0, 0] = (1.0, 0.0, 0.5)
    want = host.quantize_frame(frame)[::-1]
    host.save_screenshot(str(tmp_path / "s.ppm"), frame)
    with open(tmp_path / "s.ppm", "rb") as f:
        assert f.readline() == b"P6\n" and f.readline() == b"53 37\n" and f.readline() == b"255\n"
        assert np.array_equal(np.frombuffer(f.read(), np.uint8).reshape(37, 53, 3), want)
    host.save_screenshot(str(tmp_path / "s.png"), frame)
    try:
        from PIL import Image
    except ImportError:
        return
    im = Image.open(tmp_path / "s.png")
    assert im.size == (53, 37) and im.mode == "RGB"
    assert np.array_equal(np.asarray(im), want)
    big = rng.uniform(0, 1, (300, 400, 3)).astype(np.float32)      # several stored deflate blocks
    host.save_screenshot(str(tmp_path / "b.png"), big)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "b.png")), host.quantize_frame(big)[::-1])
