#!/usr/bin/env python
"""Device-timed numbers for every BASELINE.json config on one GPU (not the
driver's bench line; a survey of the other configs).

    python tools/bench_configs.py [--variant exact|fast] [--reps 10] [--skip-100k]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ray_tracing_b200 import host, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="exact")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--skip-100k", action="store_true")
    ap.add_argument("--kernel", default="auto", choices=["auto", "pixel", "persistent", "wavefront", "queued"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch

    variant = host.RT_VARIANT_FAST if a.variant == "fast" else host.RT_VARIANT_EXACT
    kern = {"auto": 0, "pixel": 1, "persistent": 2, "wavefront": 3, "queued": 4}[a.kernel]
    faces, sky_desc = bench.load_skybox_faces()
    r = host.Renderer(num_gpus=1)
    r.upload_skybox(faces)
    cam = host.Camera()
    out = []

    def timed(label, w, h, **kw):
        frame = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        best, st = 1e9, None
        for _ in range(a.reps):
            st = r.render_into(cam, frame.data_ptr(), w, h, stats=True, variant=variant, kernel=kern, **kw)
            best = min(best, st["render_ms"])
        # the same pose is launched a.reps times: from the third launch on the queued kernel
        # (linear-scan scenes of >= 2048 tiles) hands tiles out longest-first; `ms` is the best launch
        rec = dict(config=label, w=w, h=h, ms=best, rays=st["rays"], mrays_s=st["rays"] / best / 1e3, fps=1e3 / best, **{k: v for k, v in kw.items() if k in ("scale", "traversal")})
        out.append(rec)
        print(json.dumps(rec))

    objs = {k: host.parse_scene_string(scenes.builtin_scene_text(k)) for k in (0, 1, 2)}
    r.upload_scene(objs[0]); timed("1: scene_0 1280x720", 1280, 720, scale=1)
    r.upload_scene(objs[1]); timed("2: scene_1 1920x1080", 1920, 1080, scale=1)
    r.upload_scene(objs[2]); timed("2: scene_2 1920x1080", 1920, 1080, scale=1)
    r.upload_scene(objs[0]); timed("3: scene_0 3840x2160", 3840, 2160, scale=1)
    for s in (16, 8, 4, 2, 1):
        timed(f"4: scene_0 1920x1080 pass scale {s}", 1920, 1080, scale=s)
    # whole sweep, device resident
    frame = torch.zeros((1080, 1920, 3), dtype=torch.float32, device="cuda")
    best = 1e9
    for _ in range(a.reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, st = r.render_sweep(cam, 1920, 1080, 16, 0, ptr=frame.data_ptr(), stats=False, variant=variant, kernel=kern)
        r.synchronize(); best = min(best, time.perf_counter() - t0)
    rec = dict(config="4: scene_0 1920x1080 sweep 16->1 (5 passes, accumulate+resolve)", ms=best * 1e3, sweeps_per_s=1 / best)
    out.append(rec); print(json.dumps(rec))
    if not a.skip_100k:
        t0 = time.perf_counter(); text = scenes.synthetic_spheres_text(100000); t1 = time.perf_counter()
        big = host.parse_scene_string_large(text); t2 = time.perf_counter()
        r.upload_scene(big); r.synchronize(); t3 = time.perf_counter()
        print(json.dumps(dict(gen_s=t1 - t0, parse_s=t2 - t1, upload_and_lbvh_s=t3 - t2, objects=len(big))))
        timed("5: 100k spheres 3840x2160 (LBVH)", 3840, 2160, scale=1)
        timed("5: 100k spheres 1920x1080 (LBVH)", 1920, 1080, scale=1)
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)
    r.close()


if __name__ == "__main__":
    main()
