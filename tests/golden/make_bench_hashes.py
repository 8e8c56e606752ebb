"""Hashes (and, for config 5, tiles) of the frames bench.py renders, computed by
the UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_bench_hashes.py [--skip-tiles]

For BASELINE.json configs 1, 2a, 2b, 3 the whole frame is rendered by
oracle/_ref/libref_pixel.so (reference TUs + the per-pixel RNG key hook) with the
reference's skybox as its own loader decodes it; config 4 accumulates and
resolves the reference's five pass frames with the oracle's restatement of
main.c:394 / 476.  bench.py prints `frame_sha256` for the same frames and
compares (`frame_matches_reference`); tests/test_gpu_parity.py does too.

Config 5 (100 000 spheres at 3840x2160) is out of reach of an O(N)-per-ray
renderer as a whole frame, so eight 64x64 tiles of it are rendered pixel by
pixel through the reference's own pixel() (libref_pixel_big.so: MAX_OBJECTS
raised, nothing else changed) and stored in bench_config5_tiles.npz.
"""
import hashlib
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (config table and scene text only)
from oracle.bindings import ASSETS, Port, Ref  # noqa: E402

TILE = 64
TILE_SEED = 5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def config5_tile_origins(W, H, n=8, seed=TILE_SEED):
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(0, W - TILE)), int(rng.integers(0, H - TILE))) for _ in range(n)]


def main():
    cores = os.cpu_count() or 1
    ref = Ref("pixel")
    ref.reset_camera()
    sky = ref.load_skybox()
    port = Port()
    frames, rays = {}, {}
    for name in ("1", "2a", "2b", "3", "4"):
        cfg = bench.CONFIGS[name]
        W, H = cfg["w"], cfg["h"]
        assert ref.parse_scene_file(os.path.join(ASSETS, f"scene_{cfg['scene']}.txt"))
        T = max(t for t in range(1, min(cores, W) + 1) if W % t == 0)
        if cfg["kind"] == "frame":
            # the per-pixel key makes the frame independent of the column count T
            frame, _, r = ref.render(W, H, 1, T, 0, keyed=True)
        else:
            acc = np.zeros((H, W, 3), np.float32)
            count = np.float32(0)
            r = 0
            for p, s in enumerate((16, 8, 4, 2, 1)):
                # one column: the progressive passes of the bench use num_columns = 1
                data, _, rr = ref.render(W, H, s, 1, p, keyed=True)
                r += rr
                port.accumulate(acc, data, s)
                count = np.float32(count + np.float32(1.0) / np.float32(s * s))
            frame = port.resolve(acc, count)
        frames[name] = sha(frame)
        rays[name] = int(r)
        print(name, frames[name], r, flush=True)
    out = {"frames": frames,
           "how": "oracle/_ref/libref_pixel.so (unmodified reference + per-pixel RNG key), reference skybox decoded by the reference's load_cubemap, default pose; sha256 of the f32x3 frame, bottom row first"}

    if "--skip-tiles" not in sys.argv:
        cfg = bench.CONFIGS["5"]
        W, H = cfg["w"], cfg["h"]
        big = Ref("pixel_big")
        big.reset_camera()
        big.set_skybox(sky)
        path = "/tmp/rt_golden_spheres.txt"
        with open(path, "w") as fh:
            fh.write(bench.scene_text(cfg))
        assert big.parse_scene_file(path)
        origins = config5_tile_origins(W, H)
        aspect = np.float32(W) / np.float32(H)
        tiles = np.zeros((len(origins), TILE, TILE, 3), np.float32)

        def one(job):
            k, ty, tx = job
            x0, y0 = origins[k]
            i, j = x0 + tx, y0 + ty
            # main.c:293-296: u = 1 - (float)i/(lw-1), v = 1 - (float)j/(lh-1) (scale 1, one column)
            u = np.float32(1) - np.float32(i) / np.float32(W - 1)
            v = np.float32(1) - np.float32(j) / np.float32(H - 1)
            tiles[k, ty, tx] = big.pixel(u, v, aspect, big.pixel_key(u, v, 0))

        jobs = [(k, ty, tx) for k in range(len(origins)) for ty in range(TILE) for tx in range(TILE)]
        with ThreadPoolExecutor(cores) as ex:
            for n, _ in enumerate(ex.map(one, jobs)):
                if n % 4096 == 0:
                    print("tiles", n, "/", len(jobs), flush=True)
        np.savez_compressed(os.path.join(HERE, "bench_config5_tiles.npz"), tiles=tiles, origins=np.array(origins, np.int32),
                            size=np.array([W, H], np.int32))
        out["config5_tiles"] = {"file": "bench_config5_tiles.npz", "tile": TILE, "origins": origins, "sha256": sha(tiles),
                                "how": "oracle/_ref/libref_pixel_big.so: the reference's pixel() (O(N) scan of all 100 000 spheres per ray), one call per pixel"}
    json.dump(out, open(os.path.join(HERE, "bench_frame_hashes.json"), "w"), indent=1)
    print("written", os.path.join(HERE, "bench_frame_hashes.json"))


if __name__ == "__main__":
    main()
