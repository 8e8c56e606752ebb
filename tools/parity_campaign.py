#!/usr/bin/env python
"""Randomised parity campaign: random scenes (boxes + spheres, 1..300 objects, so
both the shared-memory scan and the LBVH), random camera poses, scales, column
counts, pass indices and kernels, CUDA path (through the C ABI) vs the oracle,
bit for bit.  Test infrastructure (uses oracle/).

    python tools/parity_campaign.py [--cases 200] [--seed 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import random_scene  # noqa: E402
from oracle.bindings import Port, procedural_skybox  # noqa: E402
from ray_tracing_b200 import host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    port = Port()
    r = host.Renderer(num_gpus=1)
    skies = [procedural_skybox(s, seed=s) for s in (16, 64, 257)]
    bad = 0
    t0 = time.time()
    rays_total = 0
    for case in range(a.cases):
        n = int(rng.choice([1, 2, 5, 9, 20, 64, 65, 100, 300]))
        objs = random_scene(n, int(rng.integers(0, 2**31)), spheres_only=bool(rng.integers(0, 4) == 0),
                            extent=float(rng.choice([3.0, 6.0, 15.0])), emissive=bool(rng.integers(0, 3) > 0))
        if rng.integers(0, 5) == 0:       # a few exactly representable, axis-aligned touching boxes (ties)
            k = min(n, 3)
            objs["type"][:k] = 0
            objs["geom"][:k, :3] = [[0, 0, 0], [1, 0, 0], [0, 1, 0]][:k]
            objs["geom"][:k, 3:] = 1.0
        sky = skies[int(rng.integers(0, len(skies)))]
        pos = rng.uniform(-12, 12, 3)
        target = rng.uniform(-3, 3, 3)
        cam = host.Camera(tuple(pos), tuple(target - pos), (0, 1, 0), float(rng.choice([30.0, 30.0, 0.6, 1.2])))
        W, H = int(rng.integers(17, 200)), int(rng.integers(9, 120))
        s = int(rng.choice([1, 1, 1, 2, 4, 8, 16, 3]))
        T = int(rng.choice([1, 1, 2, 3, 5]))
        if T > W:
            T = 1
        if W // s < 2 or H // s < 2:      # the reference divides by (W/s - 1), (H/s - 1) (main.c:293-294)
            s = 1
        p = int(rng.integers(0, 1000))
        kern = int(rng.choice([1, 2, 3]))
        r.upload_skybox(sky)
        r.upload_scene(objs)
        try:
            frame, st = r.render_frame(cam, W, H, s, num_columns=T, pass_index=p, kernel=kern)
        except host.RtError as e:
            print("case", case, "render error", e)
            bad += 1
            continue
        want, rays = port.render(port.world(objs, sky, cam.as_dict()), W, H, s, T, p)
        rays_total += rays
        same = np.array_equal(frame.view(np.uint32), want.view(np.uint32)) and st["rays"] == rays
        if not same:
            bad += 1
            d = (frame.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
            print(f"MISMATCH case {case}: n={n} {W}x{H} s={s} T={T} p={p} kernel={kern} pixels={int(d.sum())} rays {st['rays']} vs {rays}")
    print(f"{a.cases} cases, {bad} mismatches, {rays_total} rays compared, {time.time() - t0:.1f} s")
    r.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
