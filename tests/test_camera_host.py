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
