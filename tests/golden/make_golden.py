"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Everything written here is an OUTPUT of the reference compiled from
/root/reference (oracle/_ref/libref_pixel.so); the parity tests on the GPU box
compare the oracle port and the CUDA path against these files.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.bindings import ASSETS, Ref, procedural_skybox  # noqa: E402

CASES = [
    # name, scene, W, H, scale, columns, pass
    ("scene0_96x54_s1", 0, 96, 54, 1, 1, 0),
    ("scene1_96x54_s1", 1, 96, 54, 1, 1, 0),
    ("scene2_96x54_s1", 2, 96, 54, 1, 1, 0),
    ("scene0_128x72_s2_c4_p3", 0, 128, 72, 2, 4, 3),
    ("scene0_100x60_s4_c3", 0, 100, 60, 4, 3, 0),
    ("scene1_120x68_s16_c1_p1", 1, 120, 68, 16, 1, 1),
]


def main():
    ref = Ref("pixel")
    sky = procedural_skybox(64, seed=7)
    ref.set_skybox(sky)
    ref.reset_camera()
    out = {}
    meta = []
    for name, sc, W, H, s, T, p in CASES:
        assert ref.parse_scene_file(os.path.join(ASSETS, f"scene_{sc}.txt"))
        frame, _, rays = ref.render(W, H, s, T, p, keyed=True)
        out[name] = frame
        meta.append((name, sc, W, H, s, T, p, rays))
    # a moved camera pose
    ref.rotate_camera(400, 300)
    ref.rotate_camera(460, 280)
    ref.move_camera(0, 0.5)
    ref.move_camera(3, 0.5)
    cam = ref.get_camera()
    assert ref.parse_scene_file(os.path.join(ASSETS, "scene_0.txt"))
    frame, _, rays = ref.render(96, 54, 1, 1, 0, keyed=True)
    out["scene0_96x54_moved"] = frame
    out["moved_camera"] = np.concatenate([cam["pos"], cam["front"], cam["up"], [cam["fov"]]]).astype(np.float32)
    ref.reset_camera()

    # unit KATs
    out["rng_u64_state0"] = np.array(ref.rng_u64(0, 8), np.uint64)
    out["rng_f32_state0"] = ref.random_floats(0, 8)
    out["rng_dir_state0"] = ref.random_direction(0)[0]
    pts = np.array([(0, 0), (0.5, 0.5), (1, 1), (0.25, 0.75)], np.float32)
    out["camera_pxpy"] = pts
    out["camera_rays_16x9"] = np.stack([ref.camera_ray(px, py, 1280 / 720) for px, py in pts])
    rng = np.random.default_rng(5)
    dirs = rng.normal(size=(256, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    out["sky_dirs"] = dirs
    out["sky_colors"] = ref.sample_cubemap_many(dirs)
    assert ref.parse_scene_file(os.path.join(ASSETS, "scene_0.txt"))
    rays = np.concatenate([rng.uniform(-2, 8, (512, 3)), rng.normal(size=(512, 3))], axis=1).astype(np.float32)
    hit, obj = ref.trace_many(rays)
    out["trace_rays"] = rays
    out["trace_hits"] = hit
    out["trace_obj"] = obj
    # parsed scenes (field-wise; sphere union tail zeroed)
    for sc in (0, 1, 2):
        o = ref.parse_scene_file_objects(os.path.join(ASSETS, f"scene_{sc}.txt"))
        o["geom"][o["type"] == 1, 4:] = 0
        out[f"scene{sc}_objects"] = o.view(np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_vectors.txt"), "w") as f:
        f.write("# name scene W H scale columns pass rays sha256(f32 frame)\n")
        for name, sc, W, H, s, T, p, rays in meta:
            f.write(f"{name} {sc} {W} {H} {s} {T} {p} {rays} {hashlib.sha256(out[name].tobytes()).hexdigest()}\n")
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
