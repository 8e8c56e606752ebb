#!/usr/bin/env python
"""A/B harness for the LBVH kernels (development tool, GPU box only).

    python tools/lbvh_ab.py [--lib PATH] [--kernels persistent,queued] [--n 100000] [--tag X]

BASELINE config 5 (100 000 spheres) at 1080p and 4K: device time (CUDA events in
RtRenderStats, best and median of a few launches), rays, sha256 of the frame
(bit-identity across kernels / builds is the check; the parity tests pin one of
them to the oracle).  One JSON line per kernel.
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--kernels", default="persistent,queued")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tag", default="")
    ap.add_argument("--sizes", default="1920x1080,3840x2160")
    ap.add_argument("--one", action="store_true", help="render the first size with the first kernel four times and exit (for ncu: -s 3 -c 1 takes an ordered launch)")
    ap.add_argument("--counts", default=None, help="counter build (-DRT_COUNT_WALK): write nodes / tests per ray of the last size to this JSON file")
    ap.add_argument("--builder", default="sah", choices=["sah", "lbvh"], help="topology: host SAH (default) or device Morton/Karras (rt_cuda_set_bvh_builder)")
    ap.add_argument("--no-anyhit", action="store_true", help="light samples walk to their nearest hit (rt_lbvh_debug_set_anyhit(0))")
    a = ap.parse_args()
    import torch

    from ray_tracing_b200 import host, scenes
    if a.lib:
        host.LIB_PATH = os.path.abspath(a.lib)
    K = {"pixel": host.RT_KERNEL_PIXEL, "persistent": host.RT_KERNEL_PERSISTENT, "wavefront": host.RT_KERNEL_WAVEFRONT,
         "queued": host.RT_KERNEL_QUEUED, "auto": host.RT_KERNEL_AUTO}
    r = host.Renderer(num_gpus=1)
    if a.no_anyhit:
        r.set_light_anyhit(False)
    r.set_bvh_builder(host.RT_BVH_BUILDER_SAH if a.builder == "sah" else host.RT_BVH_BUILDER_LBVH)
    r.upload_skybox(scenes.procedural_skybox(256, seed=11))
    import time
    objs = host.parse_scene_string_large(scenes.synthetic_spheres_text(a.n))
    t0 = time.perf_counter()
    r.upload_scene(objs)
    r.synchronize()
    upload_ms = (time.perf_counter() - t0) * 1e3
    cam = host.Camera()
    sizes = [tuple(int(v) for v in s.split("x")) for s in a.sizes.split(",")]
    frame = torch.zeros((max(h for _, h in sizes), max(w for w, _ in sizes), 3), dtype=torch.float32, device="cuda")
    if a.one:
        w, h = sizes[0]
        for _ in range(4):          # records tile costs, builds the order, then two ordered launches
            r.render_into(cam, frame.data_ptr(), w, h, stats=True, kernel=K[a.kernels.split(",")[0]])
        r.close()
        return
    for name in a.kernels.split(","):
        out = {"kernel": name, "lib": a.lib or "default", "tag": a.tag, "n": a.n, "builder": a.builder, "upload_ms": round(upload_ms, 1)}
        for w, h in sizes:
            ts, st = [], None
            frame.fill_(-1.0)
            for i in range(a.reps + 1):
                st = r.render_into(cam, frame.data_ptr(), w, h, stats=True, kernel=K[name])
                if i >= 1:
                    ts.append(st["render_ms"])
            ts.sort()
            sha = hashlib.sha256(frame.view(-1)[: w * h * 3].cpu().numpy().tobytes()).hexdigest()[:16]
            out[f"{w}x{h}"] = dict(best=round(ts[0], 3), med=round(ts[len(ts) // 2], 3), rays=st["rays"], mrays_s=round(st["rays"] / ts[0] / 1e3, 1), sha=sha)
            if a.counts:
                nodes, tests = r.walk_counts()
                out[f"{w}x{h}"].update(nodes_per_ray=nodes / st["rays"], tests_per_ray=tests / st["rays"])
                json.dump(dict(workload=f"{a.n} spheres {w}x{h}, kernel {name}", rays=st["rays"], nodes=nodes, tests=tests,
                               nodes_per_ray=nodes / st["rays"], tests_per_ray=tests / st["rays"],
                               how="library built with -DRT_COUNT_WALK (tools/build_variant.sh count -DRT_COUNT_WALK=1), tools/lbvh_ab.py --counts"),
                          open(a.counts, "w"), indent=1)
        print(json.dumps(out), flush=True)
    r.close()


if __name__ == "__main__":
    main()
