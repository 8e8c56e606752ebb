"""Parity tests proper: the CUDA path, called through the C ABI
(libraytrace_b200.so), against the oracle on the same seeded inputs, against
the committed golden vectors (outputs of the unmodified reference), and
through size-independent properties at BASELINE.json's full sizes.

Bar (BASELINE.json north_star): the exact variant (-fmad=false) is bit-exact;
the fast variant differs by at most 1 LSB per 8-bit channel (main.c:666-670
quantisation) on >= 99.9 % of pixels.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_scene
from ray_tracing_b200 import host, scenes
from ray_tracing_b200.host import (RT_FB_U8X4, RT_KERNEL_AUTO, RT_KERNEL_PERSISTENT, RT_KERNEL_PIXEL, RT_KERNEL_QUEUED, RT_KERNEL_WAVEFRONT, RT_TRAVERSAL_LBVH,
                                   RT_TRAVERSAL_LINEAR, RT_VARIANT_EXACT, RT_VARIANT_FAST, Camera)

pytestmark = pytest.mark.gpu

KERNELS = [RT_KERNEL_PIXEL, RT_KERNEL_PERSISTENT, RT_KERNEL_WAVEFRONT, RT_KERNEL_QUEUED]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


# ------------------------------------------------------------------ probes


def test_native_library_is_loaded(renderer):
    maps = open("/proc/self/maps").read()
    assert "libraytrace_b200.so" in maps
    assert renderer.num_gpus == 1


def test_rng_matches_oracle_and_golden(renderer, port, gold):
    u, f = renderer.debug_rng(0, 8)
    assert [int(x) for x in u] == [int(x) for x in gold["rng_u64_state0"]]
    assert np.array_equal(bits(f), bits(gold["rng_f32_state0"]))
    for st in (1, 0xDEADBEEF, 2**63 + 5):
        u, f = renderer.debug_rng(st, 64)
        assert [int(x) for x in u] == port.rng_u64(st, 64)
        assert np.array_equal(bits(f), bits(port.random_floats(st, 64)))
        d = renderer.debug_random_directions(st, 1)[0]
        assert np.array_equal(bits(d), bits(port.random_direction(st)[0]))
    assert host.pixel_key(0.25, 0.75, 3) == port.pixel_key(0.25, 0.75, 3)


def test_camera_rays_match_oracle_and_golden(renderer, port, gold):
    got = renderer.debug_camera_rays(Camera(), gold["camera_pxpy"], 1280 / 720)
    assert np.array_equal(bits(got), bits(gold["camera_rays_16x9"]))
    rng = np.random.default_rng(1)
    for _ in range(10):
        cam = Camera(tuple(rng.uniform(-5, 5, 3)), tuple(rng.normal(size=3)), (0, 1, 0), float(rng.uniform(10, 80)))
        pts = rng.uniform(0, 1, (64, 2)).astype(np.float32)
        got = renderer.debug_camera_rays(cam, pts, 1.6)
        want = np.stack([port.camera_ray(px, py, 1.6, cam.as_dict()) for px, py in pts])
        assert np.array_equal(bits(got), bits(want))


def test_cubemap_matches_oracle_and_golden(renderer, port, gold, small_sky):
    renderer.upload_skybox(small_sky)
    assert np.array_equal(bits(renderer.debug_sample_cubemap(gold["sky_dirs"])), bits(gold["sky_colors"]))
    rng = np.random.default_rng(3)
    dirs = rng.normal(size=(20000, 3)).astype(np.float32)
    dirs[:100, 0] = dirs[:100, 1]
    dirs[100:200, 2] = -dirs[100:200, 0]
    dirs[210:220, :2] = 0.0
    assert np.array_equal(bits(renderer.debug_sample_cubemap(dirs)), bits(port.sample_cubemap_many(small_sky, dirs)))


def test_trace_matches_golden(renderer, gold):
    objs = np.frombuffer(gold["scene0_objects"].tobytes(), dtype=host.OBJECT_DTYPE)
    renderer.upload_scene(objs)
    for variant in (RT_VARIANT_EXACT,):
        hit, obj = renderer.debug_trace(gold["trace_rays"], variant=variant)
        assert np.array_equal(obj, gold["trace_obj"])
        assert np.array_equal(bits(hit), bits(gold["trace_hits"]))


@pytest.mark.parametrize("seed,n,spheres", [(0, 40, False), (1, 200, True), (2, 1024, False)])
def test_trace_random_scenes_linear_and_lbvh(renderer, port, seed, n, spheres):
    objs = random_scene(n, seed, spheres_only=spheres)
    renderer.upload_scene(objs)
    rng = np.random.default_rng(seed + 100)
    m = 20000
    rays = np.concatenate([rng.uniform(-8, 8, (m, 3)), rng.normal(size=(m, 3))], axis=1).astype(np.float32)
    tgt = objs["geom"][rng.integers(0, n, m // 2), :3] + rng.normal(scale=0.3, size=(m // 2, 3))
    rays[m // 2:, 3:] = tgt - rays[m // 2:, :3]
    rays[:50, 3:] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 50)] * rng.choice([-1, 1], (50, 1))
    rays[50:60, 3:] = 1e-7
    rays[60:70, 4] = 0.0
    want_hit, want_obj = port.trace_many(objs, rays)
    hit, obj = renderer.debug_trace(rays, traversal=RT_TRAVERSAL_LINEAR)
    assert np.array_equal(obj, want_obj)
    assert np.array_equal(bits(hit), bits(want_hit))
    if n > host.RT_LBVH_THRESHOLD:
        hit, obj = renderer.debug_trace(rays, traversal=RT_TRAVERSAL_LBVH)
        assert np.array_equal(obj, want_obj)
        assert np.array_equal(bits(hit), bits(want_hit))


def test_negative_zero_coordinates_take_the_plain_path(renderer, port, small_sky):
    """A box coordinate of -0 is outside the hoisted-division guard (the sign of
    a zero quotient); such scenes must still match the oracle."""
    objs = host.parse_scene_string("cube origin {-0 0 5} size {5 1 1}\ncube origin {4 -0 -0} size {1 3 3}\nsphere center {5 1 3} radius 1")
    assert np.signbit(objs[0]["geom"][0]) and np.signbit(objs[1]["geom"][1])
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    cam = Camera((0.0, 5.0, 0.0), (1.0, -1.0, 1.0), (0, 1, 0), 30.0)      # origin components exactly +0
    frame, st = renderer.render_frame(cam, 160, 90, 1)
    want, rays = port.render(port.world(objs, small_sky, cam.as_dict()), 160, 90, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want)) and st["rays"] == rays


def test_hoisted_division_is_ieee(renderer):
    """The slab test divides by a per-ray refined reciprocal (rt_device.cuh:
    div_hoisted); inside the guarded operand range it must equal IEEE a / b bit
    for bit.  ~1.2e9 operand pairs incl. all-ones / all-zeros mantissas and +0 numerators."""
    assert renderer.div_check(1, 148 * 8, 1024) == 0                       # the guarded range
    assert renderer.div_check(2, 148 * 8, 1024, -40, -38, 58, 60) == 0     # largest quotients
    assert renderer.div_check(3, 148 * 8, 1024, 38, 40, -60, -58) == 0     # smallest quotients
    assert renderer.div_check(4, 148 * 8, 1024, -1, 1, -10, 10) == 0       # typical magnitudes
    # outside the guard the sequence is NOT exact (that is why the guard exists)
    assert renderer.div_check(5, 148, 256, 10, 20, -126, -120) > 0         # denormal quotients


# ------------------------------------------------------------------ frames

GOLD_CASES = [
    ("scene0_96x54_s1", 0, 96, 54, 1, 1, 0),
    ("scene1_96x54_s1", 1, 96, 54, 1, 1, 0),
    ("scene2_96x54_s1", 2, 96, 54, 1, 1, 0),
    ("scene0_128x72_s2_c4_p3", 0, 128, 72, 2, 4, 3),
    ("scene0_100x60_s4_c3", 0, 100, 60, 4, 3, 0),
    ("scene1_120x68_s16_c1_p1", 1, 120, 68, 16, 1, 1),
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name,sc,W,H,s,T,p", GOLD_CASES)
def test_frames_equal_reference_golden(renderer, gold, small_sky, builtin_objects, kernel, name, sc, W, H, s, T, p):
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[sc])
    frame, st = renderer.render_frame(Camera(), W, H, s, num_columns=T, pass_index=p, kernel=kernel)
    assert np.array_equal(bits(frame), bits(gold[name])), name
    assert st["rays"] > 0 and st["render_ms"] > 0


def test_moved_camera_golden(renderer, gold, small_sky, builtin_objects):
    c = gold["moved_camera"]
    cam = Camera(tuple(c[0:3]), tuple(c[3:6]), tuple(c[6:9]), float(c[9]))
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    frame, _ = renderer.render_frame(cam, 96, 54, 1)
    assert np.array_equal(bits(frame), bits(gold["scene0_96x54_moved"]))


@pytest.mark.parametrize("sc", [0, 1, 2])
@pytest.mark.parametrize("kernel", KERNELS)
def test_720p_bit_exact_vs_oracle(renderer, port, real_sky, builtin_objects, sc, kernel):
    """BASELINE.json config 1 size (1280x720, scale 1, default pose)."""
    renderer.upload_skybox(real_sky)
    renderer.upload_scene(builtin_objects[sc])
    frame, st = renderer.render_frame(Camera(), 1280, 720, 1, kernel=kernel)
    want, rays = port.render(port.world(builtin_objects[sc], real_sky), 1280, 720, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays
    assert st["pixels"] == 1280 * 720


@pytest.mark.parametrize("sc", [1, 2])
def test_1080p_bit_exact_vs_oracle(renderer, port, real_sky, builtin_objects, sc):
    """BASELINE.json config 2 (scene_1 / scene_2 at 1920x1080)."""
    renderer.upload_skybox(real_sky)
    renderer.upload_scene(builtin_objects[sc])
    frame, st = renderer.render_frame(Camera(), 1920, 1080, 1)
    want, rays = port.render(port.world(builtin_objects[sc], real_sky), 1920, 1080, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays


@pytest.mark.parametrize("W,H,s,T,p", [(161, 91, 2, 3, 5), (200, 120, 8, 4, 1), (192, 108, 16, 2, 0), (1920, 1080, 16, 1, 2), (333, 77, 1, 7, 9)])
def test_scales_columns_passes_vs_oracle(renderer, port, small_sky, builtin_objects, W, H, s, T, p):
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    world = port.world(builtin_objects[0], small_sky)
    want, rays = port.render(world, W, H, s, T, p)
    for kernel in KERNELS:
        frame, st = renderer.render_frame(Camera(), W, H, s, num_columns=T, pass_index=p, kernel=kernel)
        assert np.array_equal(bits(frame), bits(want))
        assert st["rays"] == rays


def test_random_scene_random_camera(renderer, port, small_sky):
    objs = random_scene(60, 11)
    cam = Camera((7.5, 4.0, 9.0), (-0.7, -0.3, -0.8), (0, 1, 0), 30.0)
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    frame, st = renderer.render_frame(cam, 320, 200, 1, pass_index=2)
    want, rays = port.render(port.world(objs, small_sky, cam.as_dict()), 320, 200, 1, 1, 2)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays


def test_north_star_entry_point(renderer, port, small_sky, builtin_objects):
    renderer.upload_skybox(small_sky)
    frame = renderer.render_frame_simple(builtin_objects[1], Camera(), 160, 90, 2)
    want, _ = port.render(port.world(builtin_objects[1], small_sky), 160, 90, 2, 1, 0)
    assert np.array_equal(bits(frame), bits(want))
    # scene == NULL reuses the upload
    again = renderer.render_frame_simple(None, Camera(), 160, 90, 2)
    assert np.array_equal(bits(again), bits(want))


def test_edge_cases(renderer, port, small_sky):
    renderer.upload_skybox(small_sky)
    # empty scene: every pixel is sky
    empty = np.zeros(0, host.OBJECT_DTYPE)
    renderer.upload_scene(empty)
    frame, st = renderer.render_frame(Camera(), 64, 36, 1)
    want, rays = port.render(port.world(empty, small_sky), 64, 36, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want)) and st["rays"] == rays == 64 * 36
    # camera inside a box ignores that box (scene.c:67 returns tnear < 0)
    box = host.parse_scene_string("cube origin {4 4 4} size {2 2 2}\nsphere center {0 0 0} radius 1")
    renderer.upload_scene(box)
    frame, _ = renderer.render_frame(Camera(), 64, 36, 1)
    want, _ = port.render(port.world(box, small_sky), 64, 36, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want))
    # a 1024-object scene (MAX_OBJECTS) through both traversals
    objs = random_scene(1024, 5)
    renderer.upload_scene(objs)
    want, rays = port.render(port.world(objs, small_sky), 96, 54, 1, 1, 0)
    for trav in (RT_TRAVERSAL_LINEAR, RT_TRAVERSAL_LBVH):
        frame, st = renderer.render_frame(Camera(), 96, 54, 1, traversal=trav)
        assert np.array_equal(bits(frame), bits(want)) and st["rays"] == rays
    # bad arguments are reported, not aborted on
    with pytest.raises(host.RtError):
        renderer.render_frame(Camera(), 64, 36, 0)
    with pytest.raises(host.RtError):
        renderer.render_frame(Camera(), 64, 36, 2, rows=(3, 20))
    with pytest.raises(host.RtError):      # H/scale == 1: the reference's v = j/(lh-1) divides by zero
        renderer.render_frame(Camera(), 95, 13, 8)


def test_fast_variant_within_tolerance(renderer, port, real_sky, builtin_objects):
    """FMA-contracted build: <= 1 LSB per 8-bit channel on >= 99.9 % of pixels."""
    renderer.upload_skybox(real_sky)
    for sc in (0, 1, 2):
        renderer.upload_scene(builtin_objects[sc])
        exact, _ = renderer.render_frame(Camera(), 1280, 720, 1, variant=RT_VARIANT_EXACT)
        fast, _ = renderer.render_frame(Camera(), 1280, 720, 1, variant=RT_VARIANT_FAST)
        qa = host.quantize_frame(exact).astype(np.int32)
        qb = host.quantize_frame(fast).astype(np.int32)
        ok = (np.abs(qa - qb).max(axis=-1) <= 1).mean()
        assert ok >= 0.999, (sc, ok)


def test_u8x4_framebuffer(renderer, small_sky, builtin_objects):
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    f32, _ = renderer.render_frame(Camera(), 160, 90, 1)
    u8, _ = renderer.render_frame(Camera(), 160, 90, 1, fb_format=RT_FB_U8X4)
    assert np.array_equal(u8[..., :3], host.quantize_frame(f32))
    assert (u8[..., 3] == 255).all()


def test_sweep_sign_shortcut_equals_literal_path(renderer, small_sky, builtin_objects):
    """The light-sample sweep decides rd.n > 0 from the un-normalised vector when
    the sign is certain (rt_device.cuh: sample_faces_surface).  Forcing the
    literal normalise-then-dot path for every sample must give the same frame."""
    renderer.upload_skybox(small_sky)
    for sc in (0, 1):
        renderer.upload_scene(builtin_objects[sc])
        a, sa = renderer.render_frame(Camera(), 640, 360, 1, pass_index=4)
        renderer.set_sweep_threshold(1e30)
        try:
            b, sb = renderer.render_frame(Camera(), 640, 360, 1, pass_index=4)
        finally:
            renderer.set_sweep_threshold(-1.0)
        assert np.array_equal(bits(a), bits(b)) and sa["rays"] == sb["rays"]


def test_pipelined_host_readback(renderer, small_sky, builtin_objects):
    """opts.pipeline: calls return before their device->host copy finished;
    after rt_cuda_synchronize() every frame equals the synchronous call's."""
    import torch

    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    W, H = 320, 180
    want = [renderer.render_frame(Camera(), W, H, 1, pass_index=p)[0] for p in range(5)]
    bufs = [torch.zeros((H, W, 3), dtype=torch.float32, pin_memory=True) for _ in range(5)]
    for p in range(5):
        renderer.render_into(Camera(), bufs[p].data_ptr(), W, H, host=True, pipeline=1, scale=1, pass_index=p)
    renderer.synchronize()
    for p in range(5):
        assert np.array_equal(bits(bufs[p].numpy()), bits(want[p])), p


# ------------------------------------------------- accumulate / sweep (R12)


def test_progressive_sweep_vs_oracle(renderer, port, small_sky, builtin_objects):
    """BASELINE.json config 4 semantics at a small size, plus 1080p below."""
    W, H = 192, 108          # 108 % 16 = 12: rows the scale-16 pass never writes
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    world = port.world(builtin_objects[0], small_sky)
    frame, st = renderer.render_sweep(Camera(), W, H, 16, first_pass=0)
    acc = np.zeros((H, W, 3), np.float32)
    count = np.float32(0)
    rays = 0
    for p, s in enumerate((16, 8, 4, 2, 1)):
        data, r = port.render(world, W, H, s, 1, p)
        rays += r
        port.accumulate(acc, data, s)
        count = np.float32(count + np.float32(1.0) / np.float32(s * s))
    want = port.resolve(acc, count)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays
    assert renderer.accum_count() == count
    for kern in (RT_KERNEL_PIXEL, RT_KERNEL_WAVEFRONT, RT_KERNEL_QUEUED):
        other, so = renderer.render_sweep(Camera(), W, H, 16, first_pass=0, kernel=kern)
        assert np.array_equal(bits(other), bits(want)) and so["rays"] == rays
    frame, st = renderer.render_sweep(Camera(), W, H, 16, first_pass=0)
    # continuing at scale 1 keeps averaging (workers stay at scale 1, main.c:402)
    frame2, _ = renderer.render_frame(Camera(), W, H, 1, pass_index=5, accumulate=1)
    data, _ = port.render(world, W, H, 1, 1, 5)
    port.accumulate(acc, data, 1)
    count = np.float32(count + np.float32(1))
    assert np.array_equal(bits(frame2), bits(port.resolve(acc, count)))
    renderer.accum_reset()
    assert renderer.accum_count() == 0.0


def test_concurrent_sweep_equals_sequential_passes(renderer, small_sky, builtin_objects):
    """rt_cuda_render_sweep on one GPU runs the passes of a sweep side by side and folds them in
    pass order with one resolve kernel (rt_api.cu: sweep_concurrent).  Frame, accumulation
    (checked through one more accumulated pass), weight and ray count must equal the
    one-pass-after-the-other path: column counts that leave pixels uncovered (W % T != 0), sizes
    whose last rows the coarse passes never write, every init scale, the 8-bit frame format,
    a device frame on a caller's stream without statistics."""
    import torch

    renderer.upload_skybox(small_sky)
    cases = [(0, 192, 108, 16, 1), (0, 200, 113, 16, 7), (1, 161, 97, 8, 3), (2, 128, 72, 4, 1), (0, 96, 54, 2, 5), (1, 322, 182, 4, 4)]
    try:
        for scene, W, H, init, cols in cases:
            renderer.upload_scene(builtin_objects[scene])
            got = {}
            for on in (True, False):
                renderer.set_concurrent_sweep(on)
                frame, st = renderer.render_sweep(Camera(), W, H, init, first_pass=3, num_columns=cols)
                count = renderer.accum_count()
                nxt, _ = renderer.render_frame(Camera(), W, H, 1, pass_index=40, accumulate=1, num_columns=cols)
                u8, _ = renderer.render_sweep(Camera(), W, H, init, first_pass=3, num_columns=cols, fb_format=RT_FB_U8X4)
                got[on] = (frame, st["rays"], st["pixels"], count, nxt, u8)
            a, b = got[True], got[False]
            assert np.array_equal(bits(a[0]), bits(b[0])), (scene, W, H, init, cols)
            assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3]
            assert np.array_equal(bits(a[4]), bits(b[4])), "the accumulation buffers differ"
            assert np.array_equal(a[5], b[5])
        # stream-ordered: device frame, caller's stream, no statistics
        W, H = 192, 108
        renderer.upload_scene(builtin_objects[0])
        renderer.set_concurrent_sweep(True)
        want, _ = renderer.render_sweep(Camera(), W, H, 16, first_pass=0)
        dev = torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda")
        stream = torch.cuda.Stream()
        renderer.render_sweep(Camera(), W, H, 16, first_pass=0, ptr=dev.data_ptr(), stats=False, stream=stream.cuda_stream)
        stream.synchronize()
        assert np.array_equal(bits(dev.cpu().numpy()), bits(want))
    finally:
        renderer.set_concurrent_sweep(True)


def test_update_frame_refines_a_fresh_pose_in_one_call(renderer, port, small_sky, builtin_objects):
    """rt_cuda_update_frame with a negative budget renders every remaining pass of the ladder; for a
    fresh pose the passes run side by side (sweep_concurrent).  Frame, weight and the passes that
    follow must equal the reference's one-pass-at-a-time schedule (main.c:354, 402-403), with the
    concurrent ladder and without it."""
    W, H = 192, 108
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    world = port.world(builtin_objects[0], small_sky)
    def ladder(p0):
        acc = np.zeros((H, W, 3), np.float32)
        count = np.float32(0)
        for k, s in enumerate((8, 4, 2, 1)):
            data, _ = port.render(world, W, H, s, 1, p0 + k)
            port.accumulate(acc, data, s)
            count = np.float32(count + np.float32(1.0) / np.float32(s * s))
        return acc, count

    try:
        for concurrent in (True, False):
            renderer.set_concurrent_sweep(concurrent)
            renderer.set_progressive(8, 1)
            renderer.invalidate_accumulation()
            p0 = renderer.next_pass_index()
            acc, count = ladder(p0)
            frame, st = renderer.update_frame(Camera(), W, H, budget_ms=-1.0)
            assert np.array_equal(bits(frame), bits(port.resolve(acc, count))), concurrent
            assert renderer.accum_count() == count and renderer.next_pass_index() == p0 + 4
            frame, st = renderer.update_frame(Camera(), W, H, budget_ms=-1.0)      # at scale 1: exactly one more pass
            data, _ = port.render(world, W, H, 1, 1, p0 + 4)
            port.accumulate(acc, data, 1)
            assert np.array_equal(bits(frame), bits(port.resolve(acc, np.float32(count + np.float32(1))))), concurrent
            # stopped after the first pass of the ladder: the rest of it, one pass after the other
            renderer.invalidate_accumulation()
            p0 = renderer.next_pass_index()
            acc, count = ladder(p0)
            renderer.update_frame(Camera(), W, H, budget_ms=0.0)                  # the scale-8 pass
            frame, st = renderer.update_frame(Camera(), W, H, budget_ms=-1.0)
            assert np.array_equal(bits(frame), bits(port.resolve(acc, count))), concurrent
    finally:
        renderer.set_concurrent_sweep(True)
        renderer.set_progressive(8, 1)
        renderer.invalidate_accumulation()


def test_banded_host_readback_changes_nothing(renderer, small_sky, builtin_objects):
    """A synchronous call with a host frame renders the frame as row bands and copies band k
    while band k+1 renders (rt_api.cu: render_pass).  1, 4 and 8 bands must give the same
    frame, ray count and accumulation: plain and accumulated passes, uncovered columns and
    rows (W % T, H % scale), the 8-bit format, a caller's row band."""
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    W, H = 1600, 901
    cases = [dict(scale=1), dict(scale=2, num_columns=7, pass_index=3), dict(scale=1, fb_format=RT_FB_U8X4),
             dict(scale=4, rows=(128, 896)), dict(scale=1, accumulate=1, pass_index=9)]
    try:
        got = {}
        for bands in (1, 4, 8):
            renderer.set_sync_bands(-bands if bands > 1 else 1)      # negative: exactly that many, whatever the size
            renderer.accum_reset()
            out = []
            for kw in cases:
                kw = dict(kw)
                scale = kw.pop("scale")
                frame, st = renderer.render_frame(Camera(), W, H, scale, **kw)
                out.append((frame, st["rays"]))
            got[bands] = out
        for bands in (4, 8):
            for (fa, ra), (fb, rb), kw in zip(got[1], got[bands], cases):
                r0, r1 = kw.get("rows", (0, H))           # rows outside a caller's band are not the call's to define
                assert ra == rb and np.array_equal(fa[r0:r1].view(np.uint8), fb[r0:r1].view(np.uint8)), (bands, kw)
    finally:
        renderer.set_sync_bands(4)


def test_update_frame_loop_matches_reference_scheduler(renderer, port, small_sky, builtin_objects):
    """rt_cuda_update_frame / rt_cuda_invalidate_accumulation reproduce the
    workers' schedule (main.c:354, 402-408): init_scale, halving per published
    pass, restart after an invalidation; every frame is accum / count."""
    W, H = 160, 96
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    world = port.world(builtin_objects[0], small_sky)
    renderer.set_progressive(8, 4)
    renderer.invalidate_accumulation()
    acc = np.zeros((H, W, 3), np.float32)
    count = np.float32(0)
    p = renderer.next_pass_index()        # the pass counter runs on across tests and invalidations
    frames = []
    for s in (8, 4, 2, 1, 1):
        frame, st = renderer.update_frame(Camera(), W, H, 0.0)
        data, _ = port.render(world, W, H, s, 4, p)
        p += 1
        port.accumulate(acc, data, s)
        count = np.float32(count + np.float32(1.0) / np.float32(s * s))
        assert np.array_equal(bits(frame), bits(port.resolve(acc, count))), s
    # camera moved: back to init_scale, accum cleared, generation bumped
    gen = renderer.accum_generation()     # (a frame size other than the last call's had bumped it once more before)
    renderer.invalidate_accumulation()
    assert renderer.accum_generation() == gen + 1 and renderer.accum_count() == 0.0
    cam = Camera((4.5, 4.5, 4.5), (-1, -1, -1), (0, 1, 0), 30.0)
    frame, st = renderer.update_frame(cam, W, H, 0.0)
    data, _ = port.render(port.world(builtin_objects[0], small_sky, cam.as_dict()), W, H, 8, 4, p)
    acc = np.zeros((H, W, 3), np.float32)
    port.accumulate(acc, data, 8)
    assert np.array_equal(bits(frame), bits(port.resolve(acc, np.float32(1.0) / np.float32(64))))
    # a time budget buys several passes in one call
    renderer.invalidate_accumulation()
    frame, st = renderer.update_frame(Camera(), W, H, 50.0)
    assert renderer.accum_count() > 1.0 and st["kernel_launches"] >= 5
    renderer.set_progressive(8, 1)
    renderer.invalidate_accumulation()


def test_sweep_1080p_vs_oracle(renderer, port, real_sky, builtin_objects):
    W, H = 1920, 1080
    renderer.upload_skybox(real_sky)
    renderer.upload_scene(builtin_objects[0])
    world = port.world(builtin_objects[0], real_sky)
    frame, st = renderer.render_sweep(Camera(), W, H, 16, first_pass=0)
    acc = np.zeros((H, W, 3), np.float32)
    count = np.float32(0)
    for p, s in enumerate((16, 8, 4, 2, 1)):
        data, _ = port.render(world, W, H, s, 1, p)
        port.accumulate(acc, data, s)
        count = np.float32(count + np.float32(1.0) / np.float32(s * s))
    assert np.array_equal(bits(frame), bits(port.resolve(acc, count)))


# ------------------------------------------------------- bands / full sizes


def test_row_bands_equal_full_frame(renderer, small_sky, builtin_objects):
    """The multi-GPU partition rendered band by band on one
    GPU equals the single launch bit for bit."""
    from ray_tracing_b200.distributed import all_bands

    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    for W, H, s, world in ((320, 180, 1, 8), (320, 184, 4, 3)):
        full, st = renderer.render_frame(Camera(), W, H, s)
        parts = np.zeros_like(full)
        rays = 0
        for r0, r1 in all_bands(H, s, world):
            if r1 > r0:
                band, bst = renderer.render_frame(Camera(), W, H, s, rows=(r0, r1), band_only_fb=1)
                parts[r0:r1] = band
                rays += bst["rays"]
        assert np.array_equal(bits(full), bits(parts))
        assert rays == st["rays"]


def test_interleaved_row_blocks_equal_full_frame(renderer, small_sky, builtin_objects):
    """The multi-GPU work split: each rank renders the 16-row blocks b with
    b % N == rank straight into one shared frame.  Done here by one GPU playing
    every rank in turn; the union must equal the single launch, rays included."""
    import torch

    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    for W, H, s, world in ((320, 200, 1, 3), (320, 200, 4, 8), (130, 70, 16, 2), (96, 54, 3, 2)):
        full, st = renderer.render_frame(Camera(), W, H, s)
        frame = torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda")
        rays = 0
        for rank in range(world):
            bst = renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=s,
                                       interleave_count=world, interleave_index=rank)
            rays += bst["rays"]
        assert np.array_equal(bits(frame.cpu().numpy()), bits(full)), (W, H, s, world)
        assert rays == st["rays"]
    # same split with the 8-bit framebuffer (4 B/px) and a remote-style staged copy
    W, H, s, world = 320, 200, 1, 3
    full8, _ = renderer.render_frame(Camera(), W, H, s, fb_format=RT_FB_U8X4)
    frame8 = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    for rank in range(world):
        renderer.render_into(Camera(), frame8.data_ptr(), W, H, scale=s, fb_format=RT_FB_U8X4,
                             interleave_count=world, interleave_index=rank, remote_fb=int(rank != 0))
    renderer.synchronize()
    assert np.array_equal(frame8.cpu().numpy(), full8)


def test_interleaved_blocks_into_shared_host_frame(renderer, small_sky, builtin_objects):
    """Host destination + interleave: each rank copies only the row blocks it
    rendered into the (shared) host frame; untouched rows stay as they were."""
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    W, H = 320, 200
    full, st = renderer.render_frame(Camera(), W, H, 1)
    frame = np.full((H, W, 3), -1.0, np.float32)
    for rank in range(3):
        renderer.render_into(Camera(), frame.ctypes.data, W, H, host=True, scale=1, interleave_count=3, interleave_index=rank)
        from ray_tracing_b200.distributed import owned_rows

        owned = np.zeros(H, bool)
        for rr in range(rank + 1):
            owned[owned_rows(H, 1, rr, 3)] = True
        assert (frame[~owned] == -1.0).all() and (frame[owned] != -1.0).any()
    assert np.array_equal(bits(frame), bits(full))


def test_4k_properties(renderer, port, real_sky, builtin_objects):
    """BASELINE.json config 3 size (3840x2160): too large for the oracle to be
    quick, so use size-independent properties -- determinism, kernel
    equivalence, band invariance -- plus oracle parity on sampled bands."""
    W, H = 3840, 2160
    renderer.upload_skybox(real_sky)
    renderer.upload_scene(builtin_objects[0])
    a, sa = renderer.render_frame(Camera(), W, H, 1, kernel=RT_KERNEL_PERSISTENT)
    b, sb = renderer.render_frame(Camera(), W, H, 1, kernel=RT_KERNEL_PIXEL)
    assert np.array_equal(bits(a), bits(b)) and sa["rays"] == sb["rays"]
    c, _ = renderer.render_frame(Camera(), W, H, 1, kernel=RT_KERNEL_PERSISTENT)
    assert np.array_equal(bits(a), bits(c))
    # the default (queued) kernel into a frame full of sentinels: every pixel is
    # finished exactly once, whatever is left in the warps' stacks when work runs out
    import torch

    frame = torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda")
    for kern, rows in ((RT_KERNEL_QUEUED, None), (RT_KERNEL_AUTO, (16, 2144))):
        frame.fill_(-1.0)
        sq = renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, kernel=kern, rows=rows)
        got = frame.cpu().numpy()
        r0, r1 = rows or (0, H)
        assert np.array_equal(bits(got[r0:r1]), bits(a[r0:r1]))
        assert (got[:r0] == -1.0).all() and (got[r1:] == -1.0).all()
        if rows is None:
            assert sq["rays"] == sa["rays"]
    world = port.world(builtin_objects[0], real_sky)
    for r0 in (0, 1000, 2100):
        want = np.zeros((H, W, 3), np.float32)
        port.render(world, W, H, 1, 1, 0, rows=(r0, r0 + 24), out=want)
        assert np.array_equal(bits(a[r0:r0 + 24]), bits(want[r0:r0 + 24]))
    assert a.min() >= 0.0 and a.max() <= 1.0


def test_tile_schedule_changes_nothing(renderer, port, small_sky, builtin_objects):
    """The queued kernel hands a pose's tiles out longest-first, by the bounce counts
    the finest earlier launch of that pose recorded (a coarser pass seeds a finer
    one).  Scheduling only: every launch of the sequence -- natural order, recording,
    reordered -- equals the oracle, with and without interleaving, and leaves no
    pixel of a sentinel-filled frame behind."""
    import torch

    W, H = 640, 360                       # 7200 tiles: above the scheduling threshold
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    want, rays = port.render(port.world(builtin_objects[0], small_sky), W, H, 1, 1, 0)
    frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    try:
        for on in (True, False, True):
            renderer.set_tile_schedule(on)
            for launch in range(4):
                frame.fill_(-1.0)
                st = renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, kernel=RT_KERNEL_QUEUED)
                assert np.array_equal(bits(frame.cpu().numpy()), bits(want)), (on, launch)
                assert st["rays"] == rays
        # a different pass of the same pose reuses the order (costs barely move between passes)
        want5, rays5 = port.render(port.world(builtin_objects[0], small_sky), W, H, 1, 1, 5)
        frame.fill_(-1.0)
        st = renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, pass_index=5, kernel=RT_KERNEL_QUEUED)
        assert np.array_equal(bits(frame.cpu().numpy()), bits(want5)) and st["rays"] == rays5
        # a new pose rendered the way update_frame() does, coarse to fine: every pass is
        # ordered by what the pass before it saw (a fine tile takes its coarse tile's cost)
        cam2 = Camera((4.0, 4.5, 6.0), (-1.0, -0.8, -1.2), (0, 1, 0), 30.0)
        W2, H2 = 1280, 720                # scale 4: 4050 tiles, above the ordering threshold from there on
        world2 = port.world(builtin_objects[0], small_sky, cam2.as_dict())
        big = torch.empty((H2, W2, 3), dtype=torch.float32, device="cuda")
        oracle2 = [port.render(world2, W2, H2, s, 1, p_idx) for p_idx, s in enumerate((16, 8, 4, 2, 1, 1))]
        for mode in (2, True):            # 2: finer passes seeded by coarser ones (the A/B mode); True: the default
            renderer.set_tile_schedule(mode)
            for p_idx, s in enumerate((16, 8, 4, 2, 1, 1)):
                big.fill_(-1.0)
                st = renderer.render_into(cam2, big.data_ptr(), W2, H2, stats=True, scale=s, pass_index=p_idx, kernel=RT_KERNEL_QUEUED)
                w2, r2 = oracle2[p_idx]
                assert np.array_equal(bits(big.cpu().numpy()), bits(w2)) and st["rays"] == r2, (mode, s, p_idx)
        # interleaved row blocks: each rank's launches build their own order
        full = want
        for rep in range(3):
            frame.fill_(-1.0)
            for rank in range(2):
                renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, interleave_count=2, interleave_index=rank)
                renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, interleave_count=2, interleave_index=rank)
                renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, interleave_count=2, interleave_index=rank)
            assert np.array_equal(bits(frame.cpu().numpy()), bits(full))
    finally:
        renderer.set_tile_schedule(True)


def test_frames_in_flight_on_several_streams(renderer, small_sky, builtin_objects):
    """Frames of one pose issued on different streams may overlap on the GPU (bench.py keeps three in
    flight).  The pose's tile schedule is shared state: launches that record costs or build an order
    wait for every other stream, launches that only read it wait for the last writer (rt_api.cu:
    tile_schedule).  Every frame must equal the frame rendered alone, from the very first launches
    of the pose (recording, building) on, for both kernels that use the schedule."""
    import torch

    W, H = 640, 360
    renderer.upload_skybox(small_sky)
    for objs, kw in ((builtin_objects[0], {}), (host.parse_scene_string_large(scenes.synthetic_spheres_text(3000, seed=5)), {})):
        renderer.upload_scene(objs)
        want = {}
        ref = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
        renderer.set_tile_schedule(False)
        for k in range(4):
            renderer.render_into(Camera(), ref.data_ptr(), W, H, stats=True, pass_index=k, **kw)
            want[k] = ref.cpu().numpy().copy()
        renderer.set_tile_schedule(True)
        streams = [torch.cuda.Stream() for _ in range(3)]
        bufs = [torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda") for _ in range(12)]
        for cam in (Camera(), Camera((4.0, 4.5, 6.0), (-1.0, -0.8, -1.2), (0, 1, 0), 30.0), Camera()):
            for i in range(12):
                renderer.render_into(cam, bufs[i].data_ptr(), W, H, stream=streams[i % 3].cuda_stream, pass_index=i % 4, **kw)
            torch.cuda.synchronize()
            if cam.pos == Camera().pos:
                for i in range(12):
                    assert np.array_equal(bits(bufs[i].cpu().numpy()), bits(want[i % 4])), i


def test_tile_schedule_of_the_lbvh_kernel_changes_nothing(renderer, small_sky):
    """The persistent kernel over the LBVH records tile costs and hands tiles out longest-first like
    the queued kernel (a 1/8 share of BASELINE config 5 spent a quarter of its launch in the tail).
    Every launch of a pose -- image order, recording, reordered, interleaved shares -- must equal
    the unscheduled frame (which test_large_scene_lbvh_vs_oracle and the config-5 tiles pin)."""
    import torch

    W, H = 640, 360
    objs = host.parse_scene_string_large(scenes.synthetic_spheres_text(3000, seed=5))
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    try:
        renderer.set_tile_schedule(False)
        frame.fill_(-1.0)
        st0 = renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True)
        want = frame.cpu().numpy().copy()
        assert want.min() >= 0.0
        for mode in (True, 2):
            renderer.set_tile_schedule(mode)
            for launch in range(4):
                frame.fill_(-1.0)
                st = renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True)
                assert np.array_equal(bits(frame.cpu().numpy()), bits(want)) and st["rays"] == st0["rays"], (mode, launch)
            for rep in range(3):
                frame.fill_(-1.0)
                for rank in range(2):
                    renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, interleave_count=2, interleave_index=rank)
                assert np.array_equal(bits(frame.cpu().numpy()), bits(want)), (mode, rep)
    finally:
        renderer.set_tile_schedule(True)


def test_gl_presenter_fails_loudly_without_a_gl_context(renderer, small_sky, builtin_objects):
    """the CUDA-OpenGL presenter cannot be exercised on a headless box;
    what can be is that it reports the missing GL context / buffer as an error
    instead of crashing or silently rendering somewhere else."""
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    with pytest.raises(host.RtError) as e:
        renderer.gl_update_frame(Camera(), 160, 90)          # nothing registered
    assert "registered" in str(e.value)
    with pytest.raises(host.RtError) as e:
        renderer.gl_register_buffer(1, 160 * 90 * 12)        # no GL context on this thread
    assert "cudaGraphicsGLRegisterBuffer" in str(e.value)
    renderer.gl_unregister_buffer()                          # a no-op, not an error
    frame, st = renderer.render_frame(Camera(), 160, 90, 1)  # the library is still usable
    assert st["rays"] > 0


# ------------------------------------------------------------------- LBVH


def test_lbvh_equals_linear_scan_frames(renderer, small_sky):
    objs = random_scene(1024, 77, spheres_only=True, extent=10.0)
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    a, sa = renderer.render_frame(Camera(), 480, 270, 1, traversal=RT_TRAVERSAL_LINEAR)
    b, sb = renderer.render_frame(Camera(), 480, 270, 1, traversal=RT_TRAVERSAL_LBVH)
    assert np.array_equal(bits(a), bits(b)) and sa["rays"] == sb["rays"]
    for kern in (RT_KERNEL_WAVEFRONT, RT_KERNEL_QUEUED):
        c, sc = renderer.render_frame(Camera(), 480, 270, 1, traversal=RT_TRAVERSAL_LBVH, kernel=kern)
        assert np.array_equal(bits(a), bits(c)) and sa["rays"] == sc["rays"]


def test_light_samples_anyhit_changes_nothing(renderer, port, small_sky):
    """Light samples are walked in any-hit mode when the scene has exactly one emitter
    (rt_render.cu: warp_step).  Frames and ray counts must not depend on it: one emitter in the
    middle of the index range (ties go to lower AND higher indices), an emitting cube among
    cubes and spheres, two emitters (the mode must switch itself off), no emitter."""
    W, H = 320, 180
    renderer.upload_skybox(small_sky)
    cases = []
    cases.append(random_scene(1024, 77, spheres_only=True, extent=10.0))
    mixed = random_scene(600, 12, extent=7.0)
    mixed["type"][300] = 0
    mixed["geom"][300] = (-1.0, 6.0, -1.0, 2.0, 0.5, 2.0)          # the emitter (index 300) as a slab above the scene
    cases.append(mixed)
    two = random_scene(500, 13, spheres_only=True, extent=6.0)
    two["emission_power"][7] = 2.0
    two["emission_color"][7] = (1.0, 0.5, 0.25)
    cases.append(two)
    cases.append(random_scene(300, 14, spheres_only=True, extent=5.0, emissive=False))
    try:
        for objs in cases:
            renderer.upload_scene(objs)
            want, rays = port.render(port.world(objs, small_sky), W, H, 1, 1, 0)
            for on in (True, False):
                renderer.set_light_anyhit(on)
                for kern in (RT_KERNEL_AUTO, RT_KERNEL_QUEUED, RT_KERNEL_PIXEL):
                    got, st = renderer.render_frame(Camera(), W, H, 1, traversal=RT_TRAVERSAL_LBVH, kernel=kern)
                    assert np.array_equal(bits(got), bits(want)) and st["rays"] == rays, (on, kern)
    finally:
        renderer.set_light_anyhit(True)


def test_large_scene_lbvh_vs_oracle(renderer, port, small_sky):
    """Config 5 in small: 20 000 spheres in the generator's layout (objects
    beyond the shared-memory scan), far-away camera so D >> r stresses the
    fuzzy sphere test; full frame at low resolution against the oracle."""
    text = scenes.synthetic_spheres_text(20000, seed=20261017)
    objs = host.parse_scene_string_large(text)
    assert len(objs) == 20000
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    world = port.world(objs, small_sky)
    frame, st = renderer.render_frame(Camera(), 240, 136, 1)
    want, rays = port.render(world, 240, 136, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays
    far = Camera((120.0, 60.0, 140.0), (-1.0, -0.4, -1.1), (0, 1, 0), 30.0)
    frame, st = renderer.render_frame(far, 160, 90, 1, pass_index=1)
    want, rays = port.render(port.world(objs, small_sky, far.as_dict()), 160, 90, 1, 1, 1)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays


@pytest.fixture(scope="module")
def spheres_100k():
    text = scenes.synthetic_spheres_text(100000, seed=20261017)
    objs = host.parse_scene_string_large(text)
    assert len(objs) == 100000 and objs[0]["emission_power"] == 5.0
    return objs


def test_config5_100k_spheres_vs_oracle(renderer, port, small_sky, spheres_100k):
    """BASELINE.json config 5: the synthetic 100 000-sphere scene (generator seed
    20261017) through the LBVH against the O(N)-per-ray oracle: a full frame at
    reduced resolution plus a band of the 3840x2160 frame."""
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(spheres_100k)
    world = port.world(spheres_100k, small_sky)
    frame, st = renderer.render_frame(Camera(), 240, 135, 1)
    want, rays = port.render(world, 240, 135, 1, 1, 0)
    assert np.array_equal(bits(frame), bits(want))
    assert st["rays"] == rays
    W, H, r0, r1 = 3840, 2160, 1200, 1204
    band, bst = renderer.render_frame(Camera(), W, H, 1, rows=(r0, r1), band_only_fb=1)
    want = np.zeros((H, W, 3), np.float32)
    rays = port.render(world, W, H, 1, 1, 0, rows=(r0, r1), out=want)[1]
    assert np.array_equal(bits(band), bits(want[r0:r1]))
    assert bst["rays"] == rays


def test_config5_4k_properties(renderer, small_sky, spheres_100k):
    """Full 4K size of config 5 through size-independent properties: the three
    kernels agree bit for bit, the launch is deterministic, interleaved row
    blocks reassemble the frame, values stay in [0,1]."""
    import torch

    W, H = 3840, 2160
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(spheres_100k)
    a, sa = renderer.render_frame(Camera(), W, H, 1, kernel=RT_KERNEL_PERSISTENT)
    b, sb = renderer.render_frame(Camera(), W, H, 1, kernel=RT_KERNEL_WAVEFRONT)
    assert np.array_equal(bits(a), bits(b)) and sa["rays"] == sb["rays"]
    c, sc = renderer.render_frame(Camera(), W, H, 1, kernel=RT_KERNEL_PERSISTENT)
    assert np.array_equal(bits(a), bits(c))
    frame = torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda")
    rays = 0
    for rank in range(4):
        rays += renderer.render_into(Camera(), frame.data_ptr(), W, H, stats=True, scale=1, interleave_count=4, interleave_index=rank)["rays"]
    assert np.array_equal(bits(frame.cpu().numpy()), bits(a)) and rays == sa["rays"]
    assert a.min() >= 0.0 and a.max() <= 1.0
    assert sa["rays"] > 8e7


@pytest.mark.parametrize("tx,ty,shift", [(480, 540, 0), (480, 540, 1), (160, 23, 0), (97, 61, 2), (31, 33, 0)])
def test_tile_order_kernel_is_a_stable_sort_by_descending_cost(renderer, tx, ty, shift):
    """rt_api.cu: tile_order_kernel against numpy: a permutation of all tiles, costly classes first
    (costs above 15 share the top class), image order within a class; a fine tile takes the cost of
    the coarse tile covering it."""
    rng = np.random.default_rng(tx * 7 + ty + shift)
    cx, cy = (tx + (1 << shift) - 1) >> shift, (ty + (1 << shift) - 1) >> shift
    cost = rng.choice([0, 0, 0, 1, 2, 3, 5, 9, 10, 14, 15, 16, 40], size=(cy, cx)).astype(np.uint32)
    got = renderer.debug_tile_order(cost, shift, tx, ty)
    i = np.arange(tx * ty)
    cls = np.minimum(cost[(i // tx) >> shift, (i % tx) >> shift], 15)
    want = np.argsort(-cls.astype(np.int64), kind="stable").astype(np.uint32)
    assert np.array_equal(np.sort(got), i)
    assert np.array_equal(got, want)
