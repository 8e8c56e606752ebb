"""The frames bench.py times, checked against the UNMODIFIED reference.

tests/golden/bench_frame_hashes.json holds sha256 digests of BASELINE.json
configs 1, 2a, 2b, 3 (3840x2160, the whole frame) and 4 (16->1 sweep) rendered by
oracle/_ref/libref_pixel.so (reference TUs + per-pixel RNG key, skybox decoded by
the reference's loader); tests/golden/bench_config5_tiles.npz holds eight 64x64
tiles of the 3840x2160 frame of config 5 (100 000 spheres) rendered pixel by
pixel through the reference's pixel() with its O(N) scan.  Every CUDA kernel
must reproduce them bit for bit, and bench.py prints the same digests
(`frame_sha256`, `frame_matches_reference`), at every GPU count.
(reference: src/main.c:131-322, scene.c:79-190.)"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ray_tracing_b200 import host  # noqa: E402
from ray_tracing_b200.host import (RT_KERNEL_PERSISTENT, RT_KERNEL_PIXEL, RT_KERNEL_QUEUED, RT_KERNEL_WAVEFRONT, RT_VARIANT_FAST,  # noqa: E402
                                   Camera)

pytestmark = pytest.mark.gpu

KERNELS = {"pixel": RT_KERNEL_PIXEL, "persistent": RT_KERNEL_PERSISTENT, "wavefront": RT_KERNEL_WAVEFRONT, "queued": RT_KERNEL_QUEUED}
HASHES = os.path.join(GOLDEN, "bench_frame_hashes.json")
SKYDIR = os.path.join(ROOT, "oracle", "_ref", "assets", "skybox")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def stb_sky():
    try:
        return host.load_skybox_dir(SKYDIR)
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"reference skybox or stb helper not staged: {e}")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(HASHES))


def test_stb_helper_decodes_what_the_reference_loader_decodes(stb_sky, real_sky):
    """bench.py's texels (tools/librt_skybox_stb.so) == the parity tests' texels (the reference's load_cubemap)."""
    if real_sky.shape != stb_sky.shape:
        pytest.skip("reference loader not available on this box")
    assert np.array_equal(stb_sky, real_sky)


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("config", ["1", "2a", "2b", "3", "4"])
def test_bench_frames_equal_the_reference(renderer, stb_sky, gold, config, kernel):
    cfg = bench.CONFIGS[config]
    renderer.upload_skybox(stb_sky)
    renderer.upload_scene(host.parse_scene_string(bench.scene_text(cfg)))
    if cfg["kind"] == "sweep":
        frame, _ = renderer.render_sweep(Camera(), cfg["w"], cfg["h"], cfg["init_scale"], 0, kernel=KERNELS[kernel])
    else:
        frame, _ = renderer.render_frame(Camera(), cfg["w"], cfg["h"], 1, kernel=KERNELS[kernel])
    assert sha(frame) == gold["frames"][config], (config, kernel)


def test_both_builds_of_the_queued_kernel_give_the_bench_frame(renderer, stb_sky, gold):
    """Launches of 60 000 tiles and more run the queued kernel's 7-CTA build (72 registers, contrib /
    result parked in shared memory), smaller ones the 6-CTA build; the 4K bench frame from both."""
    cfg = bench.CONFIGS["3"]
    renderer.upload_skybox(stb_sky)
    renderer.upload_scene(host.parse_scene_string(bench.scene_text(cfg)))
    try:
        for dense in (False, True):
            renderer.set_queued_dense(dense)
            frame, st = renderer.render_frame(Camera(), cfg["w"], cfg["h"], 1, kernel=RT_KERNEL_QUEUED)
            assert sha(frame) == gold["frames"]["3"], dense
    finally:
        renderer.set_queued_dense(True)


@pytest.fixture(scope="module")
def spheres():
    return host.parse_scene_string_large(bench.scene_text(bench.CONFIGS["5"]))


@pytest.mark.parametrize("kernel", ["queued", "persistent", "wavefront"])
def test_config5_tiles_equal_the_reference(renderer, stb_sky, spheres, kernel):
    """Eight 64x64 tiles spread over the 3840x2160 frame of the 100 000-sphere scene:
    the LBVH walk (index tie-break) against the reference's O(N) scan."""
    g = np.load(os.path.join(GOLDEN, "bench_config5_tiles.npz"))
    W, H = (int(v) for v in g["size"])
    renderer.upload_skybox(stb_sky)
    renderer.upload_scene(spheres)
    frame, st = renderer.render_frame(Camera(), W, H, 1, kernel=KERNELS[kernel])
    assert st["rays"] > 8e7
    for k, (x0, y0) in enumerate(g["origins"]):
        got = frame[y0:y0 + 64, x0:x0 + 64]
        assert np.array_equal(got.view(np.uint32), g["tiles"][k].view(np.uint32)), (kernel, k, int(x0), int(y0))


@pytest.mark.parametrize("config", ["2a", "2b", "1", "3"])
def test_fast_variant_within_tolerance_at_bench_sizes(renderer, stb_sky, gold, config):
    """north_star tolerance of the FMA build at 1080p and 4K: <= 1 LSB per 8-bit channel
    ((uint8_t)(x*255), main.c:666-670) on >= 99.9 % of pixels.  The exact frame it is
    compared with is the reference's (digest checked here again)."""
    cfg = bench.CONFIGS[config]
    renderer.upload_skybox(stb_sky)
    renderer.upload_scene(host.parse_scene_string(bench.scene_text(cfg)))
    exact, _ = renderer.render_frame(Camera(), cfg["w"], cfg["h"], 1)
    assert sha(exact) == gold["frames"][config]
    fast, _ = renderer.render_frame(Camera(), cfg["w"], cfg["h"], 1, variant=RT_VARIANT_FAST)
    qa, qb = host.quantize_frame(exact).astype(np.int16), host.quantize_frame(fast).astype(np.int16)
    worst = np.abs(qa - qb).max(axis=-1)
    within = float((worst <= 1).mean())
    assert within >= 0.999, (config, within)


def deep_tree_scene():
    """A Karras tree 46 levels deep (tests/lbvh_sim.c reports tree_depth): one object per Morton bit
    (cells (2^b,0,0), (0,2^b,0), (0,0,2^b)) hangs the cell at the origin 30 levels down, and the
    objects in that cell have indices 1, 2, 4, ..., 2^16, a chain in the index bits of the keys.
    Everything else sits in one far-away cell."""
    from conftest import random_scene

    n = 70000
    objs = random_scene(n, seed=9, spheres_only=True, extent=1.0)
    objs["geom"][:, :3] = 1023.5
    objs["geom"][:, 3] = 0.3
    chain = [2 ** j for j in range(17)]
    objs["geom"][chain, :3] = 0.5
    free = [i for i in range(3, 200) if i not in chain]
    k = 0
    for axis in range(3):
        for b in range(10):
            c = np.full(3, 0.5)
            c[axis] = 2.0 ** b + 0.5
            objs["geom"][free[k], :3] = c
            k += 1
    objs["emission_power"] = 0
    return objs


def test_deep_tree_takes_the_local_stack_kernel(renderer, port, small_sky):
    """Trees deeper than the shared-memory stacks (RT_SMEM_STACK = 32) are walked by the
    local-memory-stack build of the persistent kernel, whatever kernel was asked for; frames
    still equal the O(N) oracle."""
    objs = deep_tree_scene()
    renderer.upload_skybox(small_sky)
    cam = Camera((9.0, 6.0, 14.0), (-0.5, -0.35, -1.0), (0, 1, 0), 30.0)
    want, rays = port.render(port.world(objs, small_sky, cam.as_dict()), 160, 90, 1, 1, 0)
    assert rays > 160 * 90
    try:
        # the depth is a property of the Karras tree (the host's SAH builder splits this scene evenly)
        for builder in (host.RT_BVH_BUILDER_LBVH, host.RT_BVH_BUILDER_SAH):
            renderer.set_bvh_builder(builder)
            renderer.upload_scene(objs)
            for kern in KERNELS.values():
                frame, st = renderer.render_frame(cam, 160, 90, 1, kernel=kern)
                assert np.array_equal(frame.view(np.uint32), want.view(np.uint32)) and st["rays"] == rays, (builder, kern)
    finally:
        renderer.set_bvh_builder(host.RT_BVH_BUILDER_SAH)


def test_both_bvh_builders_give_the_reference_frame(renderer, port, small_sky):
    """rt_cuda_set_bvh_builder: host SAH topology (default) and device Morton/Karras topology walk
    to the same hits as the reference's O(N) scan -- a mixed scene of cubes and spheres with the
    emitter in the middle of the index range, and spheres only; then objects move and the tree
    is refitted (rt_cuda_update_objects keeps whichever topology was built)."""
    from conftest import random_scene

    mixed = random_scene(4000, seed=31, extent=14.0)
    spheres = random_scene(6000, seed=32, spheres_only=True, extent=12.0)
    cam = Camera((4.0, 5.0, 3.0), (-1.0, -0.6, -0.7), (0, 1, 0), 30.0)
    renderer.upload_skybox(small_sky)
    rng = np.random.default_rng(8)
    try:
        for objs in (mixed, spheres):
            want, rays = port.render(port.world(objs, small_sky, cam.as_dict()), 192, 108, 1, 1, 0)
            moved = objs.copy()
            idx = rng.choice(len(objs), len(objs) // 50, replace=False)
            moved["geom"][idx, :3] += np.round(rng.normal(scale=3.0, size=(len(idx), 3)), 3).astype(np.float32)
            want_moved, rays_moved = port.render(port.world(moved, small_sky, cam.as_dict()), 192, 108, 1, 1, 0)
            for builder in (host.RT_BVH_BUILDER_SAH, host.RT_BVH_BUILDER_LBVH):
                renderer.set_bvh_builder(builder)
                renderer.upload_scene(objs)
                for kern in (RT_KERNEL_PERSISTENT, RT_KERNEL_QUEUED):
                    frame, st = renderer.render_frame(cam, 192, 108, 1, kernel=kern)
                    assert np.array_equal(frame.view(np.uint32), want.view(np.uint32)) and st["rays"] == rays, (builder, kern)
                renderer.update_objects(moved)
                frame, st = renderer.render_frame(cam, 192, 108, 1)
                assert np.array_equal(frame.view(np.uint32), want_moved.view(np.uint32)) and st["rays"] == rays_moved, builder
        with pytest.raises(host.RtError):
            renderer.set_bvh_builder(7)
    finally:
        renderer.set_bvh_builder(host.RT_BVH_BUILDER_SAH)


@pytest.mark.parametrize("seed,n,extent,cam", [
    (11, 5000, 20.0, ((2.0, 3.0, 2.0), (-1.0, -0.4, -0.7))),          # inside the cloud
    (12, 5000, 20.0, ((90.0, 60.0, 80.0), (-1.0, -0.6, -0.9))),       # outside: the boxes are re-padded for the distance
    (13, 3000, 5.0, ((900.0, 700.0, 800.0), (-1.0, -0.77, -0.89))),   # far away: D >> r, the fuzzy-hit regime of scene.c:110-115
])
def test_lbvh_fuzz_near_and_far_cameras(renderer, port, small_sky, seed, n, extent, cam):
    from conftest import random_scene

    objs = random_scene(n, seed=seed, spheres_only=(seed != 12), extent=extent)
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    c = Camera(cam[0], cam[1], (0, 1, 0), 30.0)
    want, rays = port.render(port.world(objs, small_sky, c.as_dict()), 200, 112, 1, 1, 0)
    for kern in (RT_KERNEL_QUEUED, RT_KERNEL_PERSISTENT):
        frame, st = renderer.render_frame(c, 200, 112, 1, kernel=kern)
        assert np.array_equal(frame.view(np.uint32), want.view(np.uint32)), (seed, kern)
        assert st["rays"] == rays


def test_refit_after_moving_objects_equals_oracle(renderer, port, small_sky):
    """rt_cuda_update_objects: 1 % of the spheres of a 20 000-sphere scene move (some
    far outside the old bounds), the LBVH is refitted, not rebuilt; frames equal the O(N) oracle on the
    moved scene and a fresh upload of it."""
    import time

    from conftest import random_scene

    objs = random_scene(20000, seed=21, spheres_only=True, extent=25.0)
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(objs)
    cam = Camera((3.0, 4.0, 2.0), (-1.0, -0.5, -0.8), (0, 1, 0), 30.0)
    before, _ = renderer.render_frame(cam, 200, 112, 1)
    rng = np.random.default_rng(5)
    moved = objs.copy()
    idx = rng.choice(len(objs), len(objs) // 100, replace=False)
    moved["geom"][idx, :3] += np.round(rng.normal(scale=4.0, size=(len(idx), 3)), 3).astype(np.float32)
    moved["geom"][idx[:5], :3] += 60.0                      # a few leave the old bounds altogether
    t0 = time.perf_counter()
    renderer.update_objects(moved)
    renderer.synchronize()
    refit_s = time.perf_counter() - t0
    got, st = renderer.render_frame(cam, 200, 112, 1)
    want, rays = port.render(port.world(moved, small_sky, cam.as_dict()), 200, 112, 1, 1, 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and st["rays"] == rays
    assert not np.array_equal(got, before)
    t0 = time.perf_counter()
    renderer.upload_scene(moved)
    renderer.synchronize()
    build_s = time.perf_counter() - t0
    again, _ = renderer.render_frame(cam, 200, 112, 1)
    assert np.array_equal(again.view(np.uint32), want.view(np.uint32))
    print(f"refit {refit_s * 1e3:.2f} ms, rebuild {build_s * 1e3:.2f} ms")
    with pytest.raises(host.RtError):
        renderer.update_objects(moved[:-1])
