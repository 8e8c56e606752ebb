#!/usr/bin/env python
"""A/B harness for kernel variants (development tool, GPU box only).

    python tools/ab_kernels.py [--lib PATH] --kernels persistent,queued [--quick]

For every kernel: sha256 of frames over a set of cases (bit-identity between
kernels / builds is the check; the persistent kernel is pinned to the oracle by
tests/test_gpu_parity.py) and device times (CUDA events inside the library's
RtRenderStats, best of a few launches after warm-up).  One JSON line per kernel.
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--kernels", default="persistent,queued")
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-hash", action="store_true")
    ap.add_argument("--no-schedule", action="store_true", help="queued kernel: keep tiles in image order")
    ap.add_argument("--overlap", action="store_true", help="also time frames dealt to 1/2/3 streams")
    ap.add_argument("--one", action="store_true", help="render scene_0 4K with the first kernel a few times and exit (for ncu)")
    args = ap.parse_args()
    import torch

    from ray_tracing_b200 import host, scenes
    if args.lib:
        host.LIB_PATH = os.path.abspath(args.lib)
    import bench

    K = {"pixel": host.RT_KERNEL_PIXEL, "persistent": host.RT_KERNEL_PERSISTENT, "wavefront": host.RT_KERNEL_WAVEFRONT,
         "queued": host.RT_KERNEL_QUEUED}
    faces, _ = bench.load_skybox_faces()
    r = host.Renderer(num_gpus=1)
    r.upload_skybox(faces)
    cam = host.Camera()
    if args.no_schedule:
        r.set_tile_schedule(False)
    frame = torch.zeros((2160, 3840, 3), dtype=torch.float32, device="cuda")

    if args.one:
        r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
        for _ in range(5):      # launch 1 natural order, 2 records tile costs, 3.. scheduled
            r.render_into(cam, frame.data_ptr(), 3840, 2160, stats=True, kernel=K[args.kernels.split(",")[0]])
        r.close()
        return

    def overlapped(w, h, n, streams, **o):
        """n frames back to back, dealt round robin to `streams` CUDA streams (two
        frame buffers); returns ms per frame between events on the first stream."""
        ss = [torch.cuda.Stream() for _ in range(streams)]
        bufs = [torch.empty((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(streams)]
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ss[0])
            for s_ in ss[1:]:
                s_.wait_event(e0)
            for i in range(n):
                k_ = i % streams
                r.render_into(cam, bufs[k_].data_ptr(), w, h, stream=ss[k_].cuda_stream, **o)
            for s_ in ss[1:]:
                ss[0].wait_stream(s_)
            e1.record(ss[0])
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        return round(best, 4), hashlib.sha256(bufs[-1].cpu().numpy().tobytes()).hexdigest()[:16]

    def sha(w, h):
        return hashlib.sha256(frame.view(-1)[: w * h * 3].cpu().numpy().tobytes()).hexdigest()[:16]

    def timed(w, h, reps, **o):
        ts = []
        st = None
        frame.fill_(-1.0)       # a pixel the launch forgets shows up in the hash
        for i in range(reps + 3):
            st = r.render_into(cam, frame.data_ptr(), w, h, stats=True, **o)
            if i >= 3:
                ts.append(st["render_ms"])
        ts.sort()
        return dict(best=round(ts[0], 4), med=round(ts[len(ts) // 2], 4), rays=st["rays"])

    for name in args.kernels.split(","):
        k = K[name]
        out = {"kernel": name, "lib": args.lib or "default", "tag": args.tag, "time": {}, "hash": {}}
        for sc in (0, 1, 2):
            r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(sc)))
            if sc == 0:
                if args.overlap:
                    for nst in (1, 2, 3):
                        out["time"]["ov%d_s0_4k" % nst] = overlapped(3840, 2160, 24, nst, kernel=k)
                        out["time"]["ov%d_s0_4k_il8" % nst] = overlapped(3840, 2160, 48, nst, kernel=k, interleave_count=8, interleave_index=3)
                        out["time"]["ov%d_s0_720p" % nst] = overlapped(1280, 720, 48, nst, kernel=k)
                out["time"]["s0_4k"] = timed(3840, 2160, args.reps, kernel=k)
                if not args.no_hash:
                    out["hash"]["s0_4k"] = sha(3840, 2160)
                out["time"]["s0_4k_il8"] = timed(3840, 2160, args.reps, kernel=k, interleave_count=8, interleave_index=3)
                out["time"]["s0_720p"] = timed(1280, 720, args.reps, kernel=k)
                out["time"]["s0_4k_fast"] = timed(3840, 2160, args.reps, kernel=k, variant=host.RT_VARIANT_FAST)
                if not args.no_hash:
                    out["hash"]["s0_4k_fast"] = sha(3840, 2160)
                    frame.fill_(-1.0)
                    r.render_into(cam, frame.data_ptr(), 1000, 562, stats=True, kernel=k, num_columns=3, pass_index=2, scale=2)
                    out["hash"]["s0_c3_s2"] = sha(1000, 562)
                    r.accum_reset()
                    frame.fill_(-1.0)
                    _, st = r.render_sweep(cam, 1920, 1080, 16, first_pass=0, ptr=frame.data_ptr(), kernel=k)
                    out["hash"]["s0_sweep16"] = sha(1920, 1080) + ":%d" % st["rays"]
                    r.accum_reset()
                    frame.fill_(-1.0)
                    r.render_into(cam, frame.data_ptr(), 1920, 1080, stats=True, kernel=k, fb_format=host.RT_FB_U8X4, scale=4)
                    out["hash"]["s0_u8_s4"] = sha(1920, 360)
            else:
                out["time"]["s%d_1080p" % sc] = timed(1920, 1080, args.reps, kernel=k)
                if not args.no_hash:
                    out["hash"]["s%d_1080p" % sc] = sha(1920, 1080)
        if not args.no_hash:
            # LBVH path: 3000 spheres in the config-5 layout, small frame
            objs = host.parse_scene_string_large(scenes.synthetic_spheres_text(3000, seed=5))
            if objs is not None:
                r.upload_scene(objs)
                out["time"]["lbvh3000_540p"] = timed(960, 540, 3, kernel=k)
                out["hash"]["lbvh3000_540p"] = sha(960, 540)
        print(json.dumps(out), flush=True)
    r.close()


if __name__ == "__main__":
    main()
