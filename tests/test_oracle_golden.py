"""The oracle port (oracle/rt_oracle.c) against the committed golden vectors,
which are outputs of the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = [
    ("scene0_96x54_s1", 0, 96, 54, 1, 1, 0),
    ("scene1_96x54_s1", 1, 96, 54, 1, 1, 0),
    ("scene2_96x54_s1", 2, 96, 54, 1, 1, 0),
    ("scene0_128x72_s2_c4_p3", 0, 128, 72, 2, 4, 3),
    ("scene0_100x60_s4_c3", 0, 100, 60, 4, 3, 0),
    ("scene1_120x68_s16_c1_p1", 1, 120, 68, 16, 1, 1),
]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_rng_known_answers(port, gold):
    # SURVEY.md R7 KAT, reproduced by the reference here
    assert port.rng_u64(0, 3) == [0x5C71580FE1214A64, 0xB8E2B01FC24294C8, 0x94A4A556CBBC9F73]
    assert port.rng_u64(0, 8) == [int(x) for x in gold["rng_u64_state0"]]
    assert np.array_equal(bits(port.random_floats(0, 8)), bits(gold["rng_f32_state0"]))
    assert np.array_equal(bits(port.random_direction(0)[0]), bits(gold["rng_dir_state0"]))
    assert float.hex(float(port.random_floats(0, 1)[0])) == "0x1.71c5600000000p-2"


def test_camera_known_answers(port, gold):
    for (px, py), want in zip(gold["camera_pxpy"], gold["camera_rays_16x9"]):
        got = port.camera_ray(px, py, 1280 / 720)
        assert np.array_equal(bits(got), bits(want))
    # SURVEY.md 8(c) camera KAT
    d = port.camera_ray(0.5, 0.5, 1280 / 720)[3:]
    assert all(float.hex(float(x)) == "-0x1.279a700000000p-1" for x in d)


def test_cubemap_known_answers(port, gold, small_sky):
    got = port.sample_cubemap_many(small_sky, gold["sky_dirs"])
    assert np.array_equal(bits(got), bits(gold["sky_colors"]))


def test_trace_known_answers(port, gold):
    objs = np.frombuffer(gold["scene0_objects"].tobytes(), dtype=__import__("oracle.bindings", fromlist=["x"]).OBJECT_DTYPE)
    hit, obj = port.trace_many(objs, gold["trace_rays"])
    assert np.array_equal(obj, gold["trace_obj"])
    assert np.array_equal(bits(hit), bits(gold["trace_hits"]))
    assert (obj >= 0).sum() > 50


@pytest.mark.parametrize("name,sc,W,H,s,T,p", CASES)
def test_frames_bit_exact(port, gold, small_sky, builtin_objects, name, sc, W, H, s, T, p):
    world = port.world(builtin_objects[sc], small_sky)
    frame, rays = port.render(world, W, H, s, T, p)
    assert np.array_equal(bits(frame), bits(gold[name])), name
    assert rays > 0


def test_moved_camera_frame(port, gold, small_sky, builtin_objects):
    c = gold["moved_camera"]
    cam = dict(pos=c[0:3], front=c[3:6], up=c[6:9], fov=float(c[9]))
    world = port.world(builtin_objects[0], small_sky, cam)
    frame, _ = port.render(world, 96, 54, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(gold["scene0_96x54_moved"]))


def test_row_bands_partition_independent(port, small_sky, builtin_objects):
    world = port.world(builtin_objects[0], small_sky)
    full, rays = port.render(world, 64, 48, 2, 1, 0)
    parts = np.zeros_like(full)
    r = 0
    for rows in ((0, 16), (16, 34), (34, 48)):
        _, k = port.render(world, 64, 48, 2, 1, 0, rows=rows, out=parts)
        r += k
    assert np.array_equal(bits(full), bits(parts))
    assert r == rays


def test_accumulate_resolve(port):
    rng = np.random.default_rng(0)
    acc = np.zeros(300, np.float32)
    count = np.float32(0)
    for s in (4, 2, 1):
        d = rng.uniform(0, 1, 300).astype(np.float32)
        before = acc.copy()
        port.accumulate(acc, d, s)
        w = np.float32(1.0) / np.float32(s * s)
        assert np.array_equal(acc, before + d * w)
        count = np.float32(count + w)
    fr = port.resolve(acc, count)
    assert np.array_equal(fr, acc * (np.float32(1.0) / count))
