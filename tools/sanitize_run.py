#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck), GPU box only:
every kernel on the linear-scan scenes and on an LBVH scene (shared-memory
traversal stacks, direction cache, finish / refill queues, parked path state, any-hit light samples), the concurrent and the
sequential progressive sweep, the banded host read-back,
the pipelined shared-frame composite with all ranks played by this process, and
a deep tree (local-memory stacks).

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_tracing_b200 import host, scenes  # noqa: E402


def main():
    r = host.Renderer(num_gpus=1)
    r.upload_skybox(scenes.procedural_skybox(64, seed=7))
    cam = host.Camera()
    for sc in (0, 1, 2):
        r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(sc)))
        for k in (host.RT_KERNEL_PIXEL, host.RT_KERNEL_PERSISTENT, host.RT_KERNEL_WAVEFRONT, host.RT_KERNEL_QUEUED):
            r.render_frame(cam, 200, 120, 1, kernel=k)
            r.render_frame(cam, 203, 121, 4, kernel=k, num_columns=3)
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
    r.render_sweep(cam, 320, 180, 16)                       # concurrent: compact cell buffers + resolve kernel
    r.render_sweep(cam, 203, 121, 8, num_columns=3, fb_format=host.RT_FB_U8X4)
    r.set_concurrent_sweep(False)
    r.render_sweep(cam, 320, 180, 16)
    r.set_concurrent_sweep(True)
    r.set_sync_bands(-4)                                    # banded host read-back, forced on a small frame
    r.render_frame(cam, 640, 512, 1)
    r.render_frame(cam, 643, 515, 2, num_columns=3, accumulate=1)
    r.set_sync_bands(4)
    ptr, _ = r.shared_frame_create(320 * 180 * 12)
    for seq in (1, 2, 3):
        for rank in range(3):
            r.render_into(cam, ptr, 320, 180, scale=1, interleave_count=3, interleave_index=rank, remote_fb=1, frame_seq=seq, frame_ack=1)
        r.shared_frame_wait(ptr, 3, seq)
        r.shared_frame_release(ptr, seq)
    r.synchronize()
    assert r.shared_frame_error(ptr) == 0
    r.shared_frame_close(ptr, owner=True)
    for builder in (host.RT_BVH_BUILDER_LBVH, host.RT_BVH_BUILDER_SAH):   # device Karras tree, then the host's SAH topology (the default)
        r.set_bvh_builder(builder)
        r.upload_scene(host.parse_scene_string_large(scenes.synthetic_spheres_text(3000, seed=5)))
        for anyhit in (True, False):                        # one emitter: light samples in any-hit mode; parked path state
            r.set_light_anyhit(anyhit)
            for k in (host.RT_KERNEL_PIXEL, host.RT_KERNEL_PERSISTENT, host.RT_KERNEL_WAVEFRONT, host.RT_KERNEL_QUEUED):
                r.render_frame(cam, 160, 90, 1, kernel=k)
        r.set_light_anyhit(True)
    moved = host.parse_scene_string_large(scenes.synthetic_spheres_text(3000, seed=5))
    moved["geom"][::50, 0] += 1.5
    r.update_objects(moved)
    r.render_frame(cam, 160, 90, 1)
    r.close()
    print("sanitize_run: done")


if __name__ == "__main__":
    main()
